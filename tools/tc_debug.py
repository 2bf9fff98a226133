"""GPU-side debugging aid: runs the tcgen05 kernel on a list of shapes, each in its own subprocess with a
hard timeout (a hung kernel must not take the whole gpurun call down), and compares with the SIMT kernel
and the fp64 oracle.  Usage: python tools/tc_debug.py [case ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (B, N, M, H, dqk, dv, Bq, causal, pad, gain)
    "tiny1tile": (1, 128, 128, 1, 128, 128, 1, False, False, 1.0),
    "tiny2tile": (1, 256, 128, 1, 128, 128, 1, False, False, 1.0),
    "kv4": (1, 256, 512, 1, 128, 128, 1, False, False, 1.0),
    "heads": (2, 256, 512, 2, 128, 128, 2, False, False, 1.0),
    "ragged": (2, 200, 333, 2, 128, 128, 2, False, False, 1.0),
    "d64": (2, 130, 300, 4, 64, 64, 2, False, False, 1.0),
    "d32_96": (2, 64, 257, 8, 32, 96, 2, False, False, 1.0),
    "d24": (2, 1, 77, 4, 24, 24, 2, False, False, 1.0),
    "d32_160": (2, 64, 257, 8, 32, 160, 2, False, False, 1.0),
    "wide256": (1, 130, 300, 2, 128, 256, 1, False, False, 1.0),
    "wide_split": (1, 300, 4096, 2, 64, 192, 1, False, True, 1.0),
    "pair_1tile": (1, 512, 128, 1, 128, 128, 1, False, False, 1.0),
    "pair_kv4": (1, 512, 512, 1, 128, 128, 1, False, False, 1.0),
    "pair_ragged": (2, 300, 700, 2, 128, 128, 2, False, False, 1.0),
    "pair_mask": (3, 400, 900, 2, 128, 128, 1, True, True, 1.0),
    "pair_d64": (2, 640, 1000, 2, 64, 128, 2, False, True, 2.0),
    "big_192": (1, 128, 256, 1, 192, 64, 1, False, False, 1.0),
    "big_322": (1, 300, 1000, 1, 322, 322, 1, False, False, 1.0),
    "big_512": (2, 200, 700, 2, 512, 512, 2, False, True, 1.0),
    "big_131": (2, 32, 784, 1, 131, 131, 1, False, False, 1.0),
    "big_causal": (2, 260, 900, 2, 256, 160, 2, True, True, 1.0),
    "big_small": (1, 40, 96, 1, 322, 322, 1, False, False, 1.0),
    "big_small2": (1, 40, 96, 1, 328, 328, 1, False, False, 1.0),
    "big_small3": (1, 40, 96, 1, 328, 256, 1, False, False, 1.0),
    "big_small4": (1, 40, 200, 1, 328, 328, 1, False, False, 1.0),
    "big_multi": (1, 128, 51200, 1, 192, 64, 1, False, False, 1.0),
    "big_long": (1, 2048, 20000, 1, 328, 328, 1, False, False, 1.0),
    "pad": (3, 40, 300, 2, 64, 64, 1, False, True, 1.0),
    "causal": (2, 100, 300, 2, 64, 64, 2, True, True, 1.0),
    "peaked": (1, 128, 4096, 2, 128, 128, 1, False, False, 6.0),
    "long": (1, 512, 16384, 8, 128, 128, 1, False, False, 1.0),
    # scores rise steadily with the key index: the exponent reference moves (accumulator rescale) at every half tile
    "ramp": (1, 256, 2048, 2, 128, 128, 1, False, False, 1.0),
    "ramp_1tile_segments": (2, 256, 640, 2, 64, 64, 2, False, False, 1.0),
    "ramp_causal": (2, 300, 1500, 2, 128, 128, 2, True, True, 1.0),
    "seg_many": (4, 256, 384, 8, 128, 128, 4, False, False, 1.0),
}


def run_case(name):
    import torch
    from perceiver_io_b200 import ops
    from gpu_util import oracle_core

    B, N, M, H, dqk, dv, Bq, causal, pad, gain = CASES[name]
    g = torch.Generator().manual_seed(0)
    q = (torch.randn(Bq, N, H * dqk, generator=g) * gain).bfloat16().cuda()
    k = torch.randn(B, M, H * dqk, generator=g).bfloat16().cuda()
    v = torch.randn(B, M, H * dv, generator=g).bfloat16().cuda()
    if name.startswith("ramp"):
        # k_j = (j * step) * u, q_n = u * |u|^-2 * sqrt(dqk)  =>  scaled score = j * step (+ noise): +12 log2 units per 64 keys
        u = torch.randn(H * dqk, generator=g)
        qn = torch.randn(Bq, N, H * dqk, generator=g) * 0.05
        kn = torch.randn(B, M, H * dqk, generator=g) * 0.05
        ramp = torch.arange(M, dtype=torch.float32)[None, :, None] * (12.0 * 0.6931 / 64.0)
        uh = u.view(H, dqk)
        uq = (uh / (uh * uh).sum(-1, keepdim=True) * dqk ** 0.5).reshape(1, 1, H * dqk)
        q = (qn + uq).bfloat16().cuda()
        k = (kn + ramp * u.view(1, 1, -1)).bfloat16().cuda()
    pm = None
    if pad:
        pm = torch.zeros(B, M, dtype=torch.bool)
        pm[0, :37] = True
        if B > 1:
            pm[1, :] = True
        if B > 2:
            pm[2, 250:] = True
        pm = pm.cuda()
    scale = dqk ** -0.5
    try:
        out = ops.attention(q, k, v, H, scale, pad_mask=pm, causal=causal, impl="tcgen05").float()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        from perceiver_io_b200 import _lib
        print("RESULT " + json.dumps({"case": name, "error": str(e)[:300], "watchdog": _lib.debug_read()[:6]}), flush=True)
        return
    ref_simt = ops.attention(q, k, v, H, scale, pad_mask=pm, causal=causal, impl="simt").float()
    res = {"case": name, "finite": bool(torch.isfinite(out).all()),
           "err_vs_simt": float((out - ref_simt).abs().max()), "ref_max": float(ref_simt.abs().max())}
    if N * M * B * H <= 2 ** 24:
        ref = oracle_core(q, k, v, H, scale, None if pm is None else pm.cpu(), causal)
        res["err_vs_oracle"] = float((out.double().cpu() - ref).abs().max())
    if res["err_vs_simt"] > 0.02 * res["ref_max"]:
        d = (out - ref_simt).abs()
        bad = (d > 0.02 * res["ref_max"])
        res["bad_frac"] = float(bad.float().mean())
        idx = bad.nonzero()[:6].tolist()
        res["bad_idx"] = idx
        res["sample_out"] = out.flatten()[:8].tolist()
        res["sample_ref"] = ref_simt.flatten()[:8].tolist()
        rows_bad = bad.any(-1)
        res["bad_rows_per_batch"] = rows_bad.sum(-1).tolist()
        cols_bad = bad.any(1).any(0)
        res["bad_cols"] = int(cols_bad.sum())
        idx = cols_bad.nonzero().flatten().tolist()
        runs, start = [], None
        for a, b2 in zip(idx, idx[1:] + [None]):
            if start is None:
                start = a
            if b2 is None or b2 != a + 1:
                runs.append((start, a))
                start = None
        res["bad_col_runs"] = runs[:12]
        r1 = 1 if out.shape[1] > 1 else 0
        res["row1_out"] = [round(float(x), 3) for x in out[0, r1, :6]] + [round(float(x), 3) for x in out[0, r1, 250:262]]
        res["row1_ref"] = [round(float(x), 3) for x in ref_simt[0, r1, :6]] + [round(float(x), 3) for x in ref_simt[0, r1, 250:262]]
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        try:
            proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], timeout=60,
                                  capture_output=True, text=True)
            lines = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                print(lines[-1][7:], flush=True)
            else:
                print(json.dumps({"case": name, "rc": proc.returncode, "stderr": proc.stderr[-1500:]}), flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": name, "hang": True}), flush=True)
            break
