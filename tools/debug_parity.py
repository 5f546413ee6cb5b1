#!/usr/bin/env python3
"""GPU-vs-oracle diagnosis helper (run under gpurun): renders the small castle with several
(spp, bounces) settings, prints ray-count differences and the worst radiance pixels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import BuiltScene  # noqa: E402
from dust_b200 import host, scenes  # noqa: E402
from dust_b200.pipeline import Context  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    ctx = Context(0)
    s = BuiltScene(scenes.castle_like(tiles=2, tile=64, height=128, props=12), po, ctx)
    lo, hi = s.bounds()
    c = (lo + hi) / 2
    eye = c + np.array((0.7, 0.5, 0.9)) * (hi - lo).max()
    W, H = 160, 96
    cam, sky = host.camera_look_at(eye, c, width=W, height=H), host.sunlight_bake()
    for spp, b in ((1, 0), (1, 1), (1, 2), (2, 2)):
        g = s.loaded.scene.trace(cam, sky, W, H, spp=spp, max_bounces=b, seed=11, hits=True, stats=True)
        o = s.oscene.render(cam, sky, W, H, spp=spp, max_bounces=b, seed=11)
        A, B = g["radiance"][..., :3].astype(np.float64), o["radiance"][..., :3].astype(np.float64)
        scale = np.maximum(np.abs(B).max(axis=-1, keepdims=True), 1e-3 * np.abs(B).mean())
        err = (np.abs(A - B) / scale).max(axis=-1)
        bad = np.argwhere(err > 1e-4)
        print(f"spp={spp} bounces={b}: max rel err {err.max():.3e}, {len(bad)} px > 1e-4",
              {k: (g['stats'][k], o['stats'][k]) for k in ("primary_rays", "bounce_rays", "shadow_rays", "hits_shaded")})
        for p in bad[:6]:
            y, x = p
            print("   px", (int(y), int(x)), "gpu", A[y, x], "oracle", B[y, x], "inst", int(o["hit_instance"][y, x]), "t", float(o["hit_t"][y, x]))
            # replay every ray of this pixel's paths through the GPU ray-batch entry point
            rays, ohits, rgba = s.oscene.debug_pixel(cam, sky, W, H, int(x), int(y), spp=spp, max_bounces=b, seed=11)
            ghits = s.loaded.scene.trace_rays(rays)
            for i in range(len(rays)):
                anyhit = bool(ohits["voxel"][i] >> 31)
                oh = (float(ohits["t"][i]), int(ohits["instance"][i]), int(ohits["leaf"][i]), int(ohits["voxel"][i] & 0xFF))
                gh = (float(ghits["t"][i]), int(ghits["instance"][i]), int(ghits["leaf"][i]), int(ghits["voxel"][i] & 0xFF))
                same = (oh[1] == 0xFFFFFFFF) == (gh[1] == 0xFFFFFFFF) if anyhit else oh == gh
                print(f"      ray {i} {'any' if anyhit else 'closest'} o={rays['o'][i]} d={rays['d'][i]} tmin={rays['tmin'][i]} tmax={rays['tmax'][i]}"
                      f" oracle={oh} gpu={gh} {'OK' if same else '<<< DIFF'}")
                if not same:
                    from dust_b200.pipeline import Scene
                    for k in sorted({oh[1], gh[1]} - {0xFFFFFFFF}):
                        mid, mat = s.ovox.instances[k]
                        one_o = po.Scene([(mat, s.omodels[mid].tree._h, s.omodels[mid].materials, s.ovox.palette)])
                        one_g = Scene(ctx, [(mat, s.loaded.trees[mid])])
                        for tmax in (rays["tmax"][i], max(oh[0], gh[0]), min(oh[0], gh[0])):
                            r1 = rays[i:i + 1].copy()
                            r1["tmax"] = tmax
                            print(f"         instance {k} alone, tmax={tmax!r}: oracle={one_o.trace_rays(r1)[0][0]} gpu={one_g.trace_rays(r1)[0]}  o2w={mat.tolist()}")


if __name__ == "__main__":
    main()
