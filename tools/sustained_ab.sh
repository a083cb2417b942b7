for l in base build/variants/libpolargpu_fuse1.so build/variants/libpolargpu_virt1.so build/variants/libpolargpu_bits8.so; do
  if [ "$l" != base ]; then export POLARGPU_LIB=$PWD/$l; else unset POLARGPU_LIB; fi
  echo "== $l"; python bench.py --steps 20 --warmup 5 --no-cpu --legs cascl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('sustained', round(d['cascl']['frames_per_s']/1e6,3), 'M frames/s frac', round(d['cascl']['roofline']['frac'],4))"
  python tools/occ_probe.py 2>/dev/null | tail -1
done
