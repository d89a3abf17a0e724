import torch
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
N=1<<30
a=torch.empty(N,dtype=torch.uint8,device='cuda'); b=torch.empty(N,dtype=torch.uint8,device='cuda')
af=a.view(torch.float32); bf=b.view(torch.float32)
ms=t(lambda: af.fill_(1.0)); print('fill   1GiB', ms, 'ms', N/ms/1e6, 'GB/s write')
ms=t(lambda: bf.copy_(af)); print('copy   1GiB', ms, 'ms', 2*N/ms/1e6, 'GB/s r+w')
ms=t(lambda: af.sum()); print('sum    1GiB', ms, 'ms', N/ms/1e6, 'GB/s read')
c=torch.empty(N//4,dtype=torch.uint8,device='cuda').view(torch.float32)
ms=t(lambda: torch.add(af[:N//16], 1.0, out=bf[:N//16])); print('add 256MiB', ms)
# write-heavy mix: read 1 part, write 4 parts
src=af[:N//16]
def mix():
    for i in range(4): bf[i*(N//16):(i+1)*(N//16)].copy_(src)
ms=t(mix); print('read 256MiB (cached) write 1GiB', ms, 'ms', (N)/ms/1e6, 'GB/s write')
