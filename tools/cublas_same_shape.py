"""Reference point for the conv kernels: cuBLAS (torch.matmul, bf16, fp32 accumulate) on the SAME GEMM shapes the
implicit-GEMM convolutions have - what a library GEMM reaches on these skinny problems (N = 64..256), next to the
8192^3 figure the roofline denominator comes from.  Not used by the product path.
  python tools/cublas_same_shape.py > profiles/rNN_cublas_same_shape.md"""
import torch

bt = torch.bfloat16


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


print("# cuBLAS bf16 GEMM on the implicit-GEMM shapes of the convolutions (torch.matmul, CUDA events, operands rotated over > L2)")
print("| layer | op | M | N | K | us | TFLOP/s |")
print("|---|---|---|---|---|---|---|")
for nimg in (16, 48):
    for name, cin, cout, k, ho in (("G7", 256, 256, 3, 32), ("G8", 256, 128, 5, 64), ("G9", 128, 64, 5, 128),
                                   ("G3", 128, 256, 4, 32), ("G2", 64, 128, 4, 64)):
        pix = nimg * ho * ho
        K = cin * k * k
        for op, (M, N, KK) in (("fprop", (pix, cout, K)), ("dgrad", (pix, cin, cout * k * k)), ("wgrad", (cout, K, pix))):
            nbuf = max(2, min(8, int(300e6 // ((M * KK + KK * N + M * N) * 2)) + 1))
            A = [torch.randn(M, KK, device="cuda", dtype=bt) for _ in range(nbuf)]
            B = [torch.randn(KK, N, device="cuda", dtype=bt) for _ in range(nbuf)]
            C = [torch.empty(M, N, device="cuda", dtype=bt) for _ in range(nbuf)]
            i = [0]

            def fn():
                j = i[0] % nbuf
                torch.matmul(A[j], B[j], out=C[j])
                i[0] += 1
            t = timeit(fn)
            print("| %s n=%d | %s | %d | %d | %d | %.1f | %.1f |" % (name, nimg, op, M, N, KK, t * 1e6, 2.0 * M * N * KK / t / 1e12), flush=True)
            del A, B, C
A = torch.randn(8192, 8192, device="cuda", dtype=bt); B = torch.randn(8192, 8192, device="cuda", dtype=bt)
t = timeit(lambda: torch.matmul(A, B), reps=10)
print("| 8192^3 | | 8192 | 8192 | 8192 | %.1f | %.1f |" % (t * 1e6, 2.0 * 8192 ** 3 / t / 1e12))
