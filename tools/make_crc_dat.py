"""Writes a CRC parity table in the reference's file format (/root/reference/CRC_6.dat): K rows of r integers, row i = coefficients
c0..c(r-1) of D^(r+i) mod g(D), UTF-16LE with byte-order mark, CRLF line ends.  python tools/make_crc_dat.py OUT [K r poly]
(defaults: 64 6 0x61 = D^6+D^5+1, which reproduces the reference's file byte for byte: tests/test_crc_table.py)"""
import sys


def crc_table_bytes(K=64, r=6, poly=0x61):
    low, mask = poly & ((1 << r) - 1), (1 << r) - 1
    cur, lines = low, []
    for _ in range(K):
        lines.append(" ".join(str((cur >> b) & 1) for b in range(r)))
        cur <<= 1
        if (cur >> r) & 1:
            cur = (cur & mask) ^ low
    return "\r\n".join(lines).encode("utf-16-le")          # no line end after the last row, as in the reference's file


if __name__ == "__main__":
    a = sys.argv[1:]
    K, r, poly = (int(a[1]), int(a[2]), int(a[3], 0)) if len(a) >= 4 else (64, 6, 0x61)
    open(a[0], "wb").write(b"\xff\xfe" + crc_table_bytes(K, r, poly))
