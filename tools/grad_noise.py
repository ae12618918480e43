"""Gradient comparison with the noise floors next to it: ours vs the mean of two reference runs, reference run-to-run,
ours run-to-run (both builds accumulate with float atomics, so every run is one draw of an order-dependent rounding).

    python tools/grad_noise.py [--cfg C1] [--ks 0.0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rade-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

from rade_gs_b200 import rawapi, scenes  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C1")
    ap.add_argument("--ks", type=float, default=0.0)
    a = ap.parse_args()
    import build_ref
    import diff_gaussian_rasterization as dgr
    ours, ref = dgr._C, build_ref.load()
    sc, coord, depth = scenes.make_config(a.cfg)
    sc = sc.to("cuda")
    g = scenes.make_upstream_grads(sc.height, sc.width, device="cuda")
    fo, fr = rawapi.forward(ours, sc, coord, depth, kernel_size=a.ks), rawapi.forward(ref, sc, coord, depth, kernel_size=a.ks)
    bo = [rawapi.backward(ours, sc, fo, g) for _ in range(2)]
    br = [rawapi.backward(ref, sc, fr, g) for _ in range(2)]
    print(f"cfg={a.cfg} ks={a.ks} lib={os.environ.get('LD_LIBRARY_PATH', '')[:60]}")
    print(f"{'tensor':10s} {'ours-vs-mean(ref)':>18s} {'ref/ref':>10s} {'ours/ours':>10s}")
    for k in rawapi.BWD_KEYS:
        if br[0][k].numel() == 0:
            continue
        refm = 0.5 * (br[0][k].double() + br[1][k].double())
        print(f"{k:10s} {rel(bo[0][k], refm):18.3e} {rel(br[0][k], br[1][k]):10.3e} {rel(bo[0][k], bo[1][k]):10.3e}")


if __name__ == "__main__":
    main()
