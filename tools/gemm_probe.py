import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
x = torch.randn(4096, 12544, device='cuda'); w = torch.randn(1024, 12544, device='cuda'); b = torch.randn(1024, device='cuda')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
for lib in ('cublas', 'cublaslt'):
    torch.backends.cuda.preferred_blas_library(lib)
    print(lib, 'linear', t(lambda: torch.nn.functional.linear(x, w, b)), 'mm', t(lambda: x @ w.t()))
wt = w.t().contiguous()
print('mm with pre-transposed weight (nn)', t(lambda: x @ wt))
xt = x.t().contiguous()
print('tn variant', t(lambda: (wt.t() @ xt)))
x2 = torch.randn(8192, 12544, device='cuda')
print('M=8192 per 4096 rows', t(lambda: torch.nn.functional.linear(x2, w, b))/2)
