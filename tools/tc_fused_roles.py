"""Developer tool (GPU box, build with LQPB_EXTRA_NVCC_FLAGS=-DLQPB_PHASE_TIMERS): which role bounds the fused block sweep.
Runs the bare inverse (lqpb_dev_tc_inverse_f32, N = 512, B = 128) with roles switched off through LQPB_FU_DBG (1 staging does
no work, 2 no MMAs, 4 epilogue does no work; the results are garbage, only the timeline counts).  One child per setting."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child():
    import torch
    from lqp_py_b200 import _abi
    L = _abi.lib()
    N, B = int(os.environ.get("N", 512)), 128
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    Lm = torch.randn(B, N, N, device=dev)
    A = (Lm.transpose(1, 2) @ Lm / N + 1.2 * torch.eye(N, device=dev)).contiguous()
    out = torch.empty_like(A)
    work = torch.empty(L.lqpb_dev_tc_inverse_work_bytes(B, N), dtype=torch.uint8, device=dev)
    buf = (C.c_ulonglong * 512)()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(4):
        if rep == 3:
            L.lqpb_debug_fu_ns(buf, 1)
            ev[0].record()
        rc = L.lqpb_dev_tc_inverse_f32(B, N, _abi.ptr(A), _abi.ptr(out), _abi.ptr(work), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _abi.check(rc, "dev_tc_inverse")
    ev[1].record()
    torch.cuda.synchronize()
    L.lqpb_debug_fu_ns(buf, 0)
    nb = N // 128
    t0 = prev = buf[0]
    parts = []
    for k in range(nb):
        row = []
        for j in range(3):
            t = buf[1 + 3 * k + j]
            if t:
                row.append(f"{(t - prev) / 1e3:6.1f}")
                prev = t
        parts.append("/".join(row))
    print(f"dbg={os.environ.get('LQPB_FU_DBG', '0')}: call {ev[0].elapsed_time(ev[1]) * 1e3:7.1f} us  sweep {(prev - t0) / 1e3:7.1f} us  "
          f"[pivot/PANEL/TRAIL per step: {'  '.join(parts)}]  waits us: S-free {buf[32] / 8e3:.0f}  M-full {buf[33] / 1e3:.0f}  "
          f"M-acc {buf[34] / 1e3:.0f}  E-acc {buf[35] / 4e3:.0f}")

    if os.environ.get("SLABS"):
        t00 = buf[64]
        for mode, nm, n in ((0, "PANEL", 12), (1, "TRAIL", 24)):
            print(nm + " slabs of step 0, staging thread 0 (us): start | +wait-free | +split/store | +group barrier | (MMA issue until next start)")
            for L in range(n):
                b0 = 64 + (mode * 24 + L) * 4
                t = [buf[b0 + i] for i in range(4)]
                nxt = buf[b0 + 4] if L + 1 < n else 0
                print(f"   slab {L:2d}: {(t[0] - t00) / 1e3:7.2f} | {(t[1] - t[0]) / 1e3:5.2f} | {(t[2] - t[1]) / 1e3:5.2f} | {(t[3] - t[2]) / 1e3:5.2f} | {((nxt - t[3]) / 1e3) if nxt else 0:5.2f}")

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child()
    else:
        for dbg in sys.argv[1:] or ["0", "1", "2", "4", "3", "5", "6", "7"]:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, LQPB_FU_DBG=dbg),
                               capture_output=True, text=True, timeout=120)
            print(r.stdout.strip() or r.stderr[-600:], flush=True)
