"""Sweep (n_per_item, ksplit) overrides per decoder layer shape; prints the median kernel time of each valid combination
and what the built-in chooser picks.  Run on the GPU box:  python tools/sweep_igemm.py [batch]"""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(n_force, k_force, batch):
    import torch
    from terrain_diffusion_b200 import _lib as L
    from terrain_diffusion_b200.layout import pack_weight_segments, to_nc8hw8
    from tools.bench_igemm import SHAPES
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lib = L.lib()
    for name, segs, cout, res, count in SHAPES:
        n = batch
        ch = (C.c_int32 * 3)(*[c for c, _ in segs], *([0] * (3 - len(segs))))
        tp = (C.c_int32 * 3)(*[t for _, t in segs], *([0] * (3 - len(segs))))
        n_item = L.igemm_choose_n(cout, n, res, res, segs)
        plan = (C.c_int32 * 4)()
        lib.tdx_debug_igemm_plan(cout, n, res, res, ch, tp, len(segs), n_item, 0, plan)
        if (n_force and plan[0] != n_force) or (k_force and plan[1] != k_force):
            continue
        acts = [to_nc8hw8(torch.randn(n, c, res, res, device=dev)) for c, _ in segs]
        wts = [torch.randn(cout, c, 3 if t == 9 else 1, 3 if t == 9 else 1, device=dev) * 0.02 for c, t in segs]
        b = pack_weight_segments(wts, n_item)
        out = torch.empty(n, cout // 8, res, res, 8, dtype=torch.bfloat16, device=dev)
        cvec = torch.ones(n, cout, device=dev)
        d = L.TdxIgemmDesc()
        for i, (c, t) in enumerate(segs):
            d.a_ptr[i] = acts[i].data_ptr(); d.a_channels[i] = c; d.a_taps[i] = t
        d.n_seg = len(segs); d.b_packed = b.data_ptr(); d.c_out = cout; d.n_per_item = n_item
        d.n_img, d.height, d.width = n, res, res
        d.epi_flags = L.EPI_EMB_SILU; d.cvec = cvec.data_ptr()
        d.out[0].ptr = out.data_ptr(); d.out[0].kind = L.OUT_RAW; d.out[0].scale = 1.0
        stream = L.current_stream_ptr()
        for _ in range(3):
            L.check(lib.tdx_igemm_run(C.byref(d), stream))
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):   # back-to-back dependent launches, as inside the forward graph
                L.check(lib.tdx_igemm_run(C.byref(d), stream))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 20)
        ts.sort()
        print(f"R|{name}|{plan[0]}|{plan[1]}|{ts[len(ts)//2]*1e3:.2f}|{ts[0]*1e3:.2f}", flush=True)


def main():
    if len(sys.argv) > 3:
        child(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
        return
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    res = {}
    for nf in (0, 64, 128, 192, 256):
        for kf in (0, 1, 2, 3, 4, 6, 8):
            if (nf == 0) != (kf == 0):
                continue
            env = dict(os.environ)
            if nf:
                env["TDX_IGEMM_N"] = str(nf)
                env["TDX_IGEMM_KSPLIT"] = str(kf)
            r = subprocess.run([sys.executable, __file__, str(nf), str(kf), str(batch)], env=env, capture_output=True,
                               text=True, timeout=300)
            if r.returncode:
                print("child failed", nf, kf, r.stderr[-400:])
            for line in r.stdout.splitlines():
                if line.startswith("R|"):
                    _, name, n, k, med, mn = line.split("|")
                    res.setdefault(name, []).append((float(med), int(n), int(k), nf == 0))
    for name, rows in res.items():
        auto = [r for r in rows if r[3]]
        rows = sorted(set(r[:3] for r in rows if not r[3]))
        a = f"auto N={auto[0][1]} ks={auto[0][2]} {auto[0][0]:.1f}us" if auto else "auto ?"
        print(f"{name:34s} {a:28s} | " + "  ".join(f"N{n}k{k}:{t:.1f}" for t, n, k in rows[:8]))


if __name__ == "__main__":
    main()
