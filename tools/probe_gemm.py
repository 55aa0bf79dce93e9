"""GPU probe: tcgen05 GEMM (K-major fwd, MN-major wgrad) vs torch matmul. Dev tool, not a test."""
import ctypes, os, sys, torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "lightning-asr_b200", "liblasr_b200.so"))
lib.lasr_strerror.restype = ctypes.c_char_p
vp, ci = ctypes.c_void_p, ctypes.c_int
lib.lasr_pwconv_fwd.argtypes = [vp, vp, vp, vp, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, vp]
lib.lasr_pwconv_wgrad.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
def P(t): return None if t is None else t.data_ptr()
def S(): return torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
dev = "cuda"
ok = True
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
for dtype, code in ((torch.bfloat16, 1), (torch.float32, 0)):
    for (M, K, N, T) in [(1000, 256, 256, 250), (2004, 64, 256, 501), (3000, 336, 512, 1000), (777, 512, 512, 777),
                         (1501, 512, 1024, 1501), (1300, 1024, 32, 650), (900, 1024, 4336, 300), (640, 256, 64, 320),
                         (25632, 256, 256, 801)]:
        x = torch.randn(M, K, device=dev).to(dtype)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dtype)
        bias = torch.randn(N, device=dev)
        nb = M // T
        lengths = torch.randint(T // 2, T + 1, (nb,), device=dev, dtype=torch.int32)
        y = torch.full((M, N), 7.0, device=dev, dtype=dtype)
        groups = ((M + 127) // 128) * 4
        stats = torch.zeros(groups, 2, N, device=dev)
        for use_bias, use_mask in ((False, True), (True, False)):
            rc = lib.lasr_pwconv_fwd(P(x), P(w), P(y), P(bias) if use_bias else None, P(lengths) if use_mask else None, T,
                                     P(stats), M, K, N, K, K, N, code, S())
            torch.cuda.synchronize()
            if rc != 0:
                print("FAIL rc", rc, lib.lasr_strerror(rc)); ok = False; continue
            ref = x.float() @ w.float().t()
            if use_bias: ref = ref + bias
            if use_mask:
                t = torch.arange(M, device=dev) % T
                n = torch.arange(M, device=dev) // T
                ref = ref * (t < lengths[n.clamp_max(nb - 1)]).unsqueeze(1)
                if nb * T < M: pass
            e = rel(y.float(), ref)
            s_ref = ref.sum(0); q_ref = (ref * ref).sum(0)
            es = rel(stats[:, 0].sum(0), s_ref) if s_ref.norm() > 0 else 0
            eq = rel(stats[:, 1].sum(0), q_ref)
            tol = 1e-2 if code == 1 else 1e-5
            flag = "ok" if (e < tol and eq < tol) else "BAD"
            if flag == "BAD": ok = False
            print(f"fwd {dtype} M{M} K{K} N{N} bias{use_bias} mask{use_mask}: rel {e:.2e} stat-sum {es:.2e} stat-sq {eq:.2e} {flag}")
        # wgrad
        dy = torch.randn(M, N, device=dev).to(dtype)
        dw = torch.zeros(N, K, device=dev)
        rc = lib.lasr_pwconv_wgrad(P(dy), P(x), P(dw), M, K, N, N, K, K, code, S())
        torch.cuda.synchronize()
        if rc != 0:
            print("FAIL wgrad rc", rc, lib.lasr_strerror(rc)); ok = False; continue
        ref = dy.float().t() @ x.float()
        e = rel(dw, ref)
        flag = "ok" if e < (1e-2 if code == 1 else 1e-4) else "BAD"
        if flag == "BAD": ok = False
        print(f"wgrad {dtype} M{M} K{K} N{N}: rel {e:.2e} {flag}")
# timing of the big one
M, K, N = 25632, 256, 256
x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16(); y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(((M + 127) // 128) * 4, 2, N, device=dev)
for (M, K, N) in [(25632, 256, 256), (25632, 512, 512), (25632, 512, 1024)]:
    x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16(); y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    dyy = torch.randn(M, N, device=dev).bfloat16(); dw = torch.zeros(N, K, device=dev)
    stats = torch.zeros(((M + 127) // 128) * 4, 2, N, device=dev)
    for name, fn in (("fwd", lambda: lib.lasr_pwconv_fwd(P(x), P(w), P(y), None, None, 0, P(stats), M, K, N, K, K, N, 1, S())),
                     ("wgrad", lambda: lib.lasr_pwconv_wgrad(P(dyy), P(x), P(dw), M, K, N, N, K, K, 1, S())),
                     ("torch", lambda: torch.matmul(x, w.t(), out=y))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"time {name} M{M} K{K} N{N}: {ms*1e3:.1f} us  {2*M*K*N/ms/1e9:.1f} TFLOP/s  {(M*K+M*N)*2/ms/1e6:.0f} GB/s")
print("ALL OK" if ok else "SOME BAD")
sys.exit(0 if ok else 1)
