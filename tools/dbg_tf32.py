import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmvae_b200 import ops
torch.manual_seed(0)
for (M, N, K) in [(128, 128, 32), (128, 128, 8), (128, 128, 64), (128, 128, 256)]:
    A = torch.ones(M, K).cuda(); Bm = torch.ones(N, K).cuda()
    C = torch.full((M, N), -7.0).cuda()
    ops.gemm(A, 0, Bm, 0, M, N, K, C32=C, tf32=True)
    torch.cuda.synchronize()
    print("ones", M, N, K, C[0, :4].tolist(), C[64, 100].item(), float(C.abs().max()))
    A = torch.arange(M * K, dtype=torch.float32).view(M, K).cuda() / 100; Bm = torch.eye(N, K).cuda()
    ops.gemm(A, 0, Bm, 0, M, N, K, C32=C, tf32=True)
    torch.cuda.synchronize()
    print("eye ", C[1, :6].tolist(), (A @ Bm.t())[1, :6].tolist())
    A16 = torch.ones(M, 64).bfloat16().cuda(); B16 = torch.ones(N, 64).bfloat16().cuda()
    ops.gemm(A16, 0, B16, 0, M, N, 64, C32=C)
    print("bf16", C[0, :4].tolist())
