import torch, time
dev=torch.device('cuda:0')
for mb in (12, 55, 256):
    n=mb*1024*1024//8
    h=torch.empty(n,dtype=torch.float64).pin_memory(); d=torch.empty(n,dtype=torch.float64,device=dev)
    h2=torch.empty(n,dtype=torch.float64).pin_memory(); d2=torch.ones(n,dtype=torch.float64,device=dev)
    s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
    for mode in ('h2d','d2h','both'):
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(20):
            if mode in ('h2d','both'):
                with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
            if mode in ('d2h','both'):
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/20
        print(mb,'MB',mode, '%.1f GB/s per direction' % (mb/1024/dt))
