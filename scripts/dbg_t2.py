import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from get_b200 import ops, _lib
M, N, K = 21600, 300, int(sys.argv[1])
dev = "cuda"
a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * K ** -0.5; out = torch.empty(M, N, device=dev)
x = torch.randn(M, N, device=dev); o1 = torch.empty(M, N, device=dev); b = torch.randn(N, device=dev)
mode = sys.argv[2]
for i in range(2):
    if mode == "store":
        ops.gemm([(a, w)], out, tc=True)
    else:
        ops.gemm([(a, w)], out, epilogue=_lib.EPI_SIGMOID, bias0=b, bias1=b, aux0=x, out1=o1, tc=True)
    torch.cuda.synchronize()
