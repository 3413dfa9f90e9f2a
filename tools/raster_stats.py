"""Instrumentation run: statistics of the compositing cull test (needs a -DRS_RASTER_STATS build of the library at
tools/_tmp/librigidsplat_stats.so: `bash tools/build_stats_lib.sh`).  Not part of the product or the tests."""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
_lib = importlib.import_module("3dgs_rigidbody_b200._lib")
_lib.LIB_PATH = os.path.join(ROOT, "tools", "_tmp", "librigidsplat_stats.so")
rs = importlib.import_module("3dgs_rigidbody_b200")
lib = _lib.load()
lib.rs_raster_stats.argtypes = [ctypes.c_void_p]
dev = "cuda:0"
sc = bench.make_domino_scene(device=dev)
fr = rs.FrameRenderer(sc["means"], sc["quats"], sc["scales"], sc["opacities"], sc["colors"], bench.WIDTH, bench.HEIGHT,
                      cluster_ids=sc["cluster_ids"], body_centers=sc["body_centers"], max_isects=24_000_000)
out = (ctypes.c_ulonglong * 8)()
for f in (0, 120, 230):
    bq, bt = bench.domino_poses(20, frame=f, device=dev, centers=sc["body_centers"])
    fr.render(sc["viewmats"], sc["Ks"], bq, bt); torch.cuda.synchronize()
    lib.rs_raster_stats(out)  # reset
    fr.render(sc["viewmats"], sc["Ks"], bq, bt); torch.cuda.synchronize()
    lib.rs_raster_stats(out)
    it, nz, passing, active, chunks = out[0], out[1], out[2], out[3], out[4]
    M = fr.n_isects()
    m = fr.meta()
    off = m["isect_offsets"].reshape(-1).long(); last = m["last_ids"].reshape(1080, 1920)
    # consumed fraction of the lists: per tile, the furthest isect any pixel blended
    print(f"frame {f}: M={M} warp-iterations={it} ({it/M:.2f} per isect of 8 possible) with>=1 passing lane={nz} ({nz/it:.2%}) "
          f"passing lanes/iter={passing/it:.1f} active lanes/iter={active/it:.1f} chunk tests={chunks} (x32 splat tests; {chunks*32/(M*8):.2%} of M*8)")
