import torch, time
x = torch.empty(1<<28, dtype=torch.float32).pin_memory()
y = torch.empty(1<<28, dtype=torch.float32, device="cuda")
for sz in (1<<24, 1<<26, 1<<28):
    for _ in range(2): y[:sz].copy_(x[:sz], non_blocking=True)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(5): y[:sz].copy_(x[:sz], non_blocking=True)
    torch.cuda.synchronize(); dt=(time.time()-t)/5
    print("H2D %d MB: %.1f GB/s" % (sz*4>>20, sz*4/dt/1e9))
z = torch.empty(1<<26, dtype=torch.float32).pin_memory()
torch.cuda.synchronize(); t=time.time()
for _ in range(5): z.copy_(y[:1<<26], non_blocking=True)
torch.cuda.synchronize(); print("D2H 256 MB: %.1f GB/s" % ((1<<28)/((time.time()-t)/5)/1e9))
