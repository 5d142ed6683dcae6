"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference
through oracle/ref_shim.py) in the build container.  The vectors are committed; this script is the
record of how they were made.  cv2 / sklearn versions are stored inside each file.

    python scripts/make_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tests.util import blobs, random_flow, synth_pair  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def versions():
    import cv2
    import sklearn
    return dict(cv2=cv2.__version__, sklearn=sklearn.__version__, numpy=np.__version__)


def pipeline_goldens(v):
    """8. the YAML / TIFF pipeline: the unmodified reference's main() (__main__.py:624-642 -> run_opt_flow_reg :534-609 ->
    register_and_save_ofreg_imgs :320-437) on the miniature datasets of tests/pipeline_data.py, its TIFF I/O served by the
    tifffile stub of oracle/ref_shim.py."""
    import tempfile
    from tests import pipeline_data as pd
    m = ref_shim.load_pipeline()
    for layout in ("per_image", "stack"):
        tmp = tempfile.mkdtemp()
        cfg, out_dir = pd.write_inputs(tmp, layout)
        buf, old = io.StringIO(), sys.argv
        sys.argv = ["microaligner", cfg]
        try:
            with contextlib.redirect_stdout(buf):
                m.main()
        finally:
            sys.argv = old
        pixels, desc = pd.read_result(out_dir)
        np.savez_compressed(os.path.join(OUT, f"pipeline_{layout}.npz"), pixels=pixels, description=desc,
                            stdout=buf.getvalue().replace(tmp, "<TMP>"), **v)


def main():
    if "--only-pipeline" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        pipeline_goldens(versions())
        for f in sorted(os.listdir(OUT)):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
        return
    mod = ref_shim.load()
    import importlib
    fc = importlib.import_module("microaligner.optflow_reg.flow_calc")
    ofr = importlib.import_module("microaligner.optflow_reg.optflow_registrator")
    sim = importlib.import_module("microaligner.shared_modules.similarity_scoring")
    os.makedirs(OUT, exist_ok=True)
    v = versions()

    # 1. one Farneback tile through the reference's own wrapper (flow_calc.py:30-47)
    ref, mov = synth_pair(240, 256, 0, np.uint16)
    flow = fc.farneback(mov, ref, 0, 99, 3)
    ref8, mov8 = synth_pair(200, 168, 1, np.uint8)
    flow8 = fc.farneback(mov8, ref8, 0, 39, 2)
    np.savez_compressed(os.path.join(OUT, "farneback_tile.npz"), ref=ref, mov=mov, flow=flow, win=99, iters=3,
                        ref8=ref8, mov8=mov8, flow8=flow8, win8=39, iters8=2, **v)

    # 2. TileFlowCalc tiled branch (flow_calc.py:59-79)
    ref, mov = synth_pair(300, 330, 2, np.uint16)
    t = fc.TileFlowCalc()
    t.ref_img, t.mov_img, t.tile_size, t.overlap, t.num_iter, t.win_size = ref, mov, 120, 20, 2, 19
    np.savez_compressed(os.path.join(OUT, "tileflow.npz"), ref=ref, mov=mov, flow=t.calc_flow(), T=120, ov=20, iters=2, win=19, **v)

    # 3. Warper (warper.py:37-76), u8 and u16
    rng = np.random.default_rng(3)
    img16 = rng.integers(0, 65536, (211, 277)).astype(np.uint16)
    img8 = rng.integers(0, 256, (211, 277)).astype(np.uint8)
    fl = random_flow(211, 277, 4, mag=6.0)
    fl[5, 5] = [1e6, -3.0]
    outs = {}
    for name, img in (("u16", img16), ("u8", img8)):
        w = mod.Warper()
        w.tile_size, w.overlap = 100, 20
        w.image, w.flow = img, fl.copy()
        outs[name] = w.warp()
    np.savez_compressed(os.path.join(OUT, "warper.npz"), img16=img16, img8=img8, flow=fl, out16=outs["u16"], out8=outs["u8"],
                        T=100, ov=20, **v)

    # 4. dog() (optflow_registrator.py:249-274) and mi_tiled (similarity_scoring.py:27-50)
    r = mod.OptFlowRegistrator()
    ref, mov = synth_pair(300, 262, 5, np.uint16)
    bl = blobs(300, 262, 6, np.uint16)
    d_ref, d_mov, d_bl = r.dog(ref, True), r.dog(mov, True), r.dog(bl, True)
    mi_whole = sim.mi_tiled(d_ref, d_mov, 1000)
    mi_chunks = sim.mi_tiled(d_ref, d_mov, 100)
    np.savez_compressed(os.path.join(OUT, "dog_nmi.npz"), ref=ref, mov=mov, blobs=bl, d_ref=d_ref, d_mov=d_mov, d_blobs=d_bl,
                        mi_whole=mi_whole, mi_chunks=mi_chunks, chunk_T=100, **v)

    # 5. merge (optflow_registrator.py:37-47, 217-240)
    f1, f2 = random_flow(230, 310, 7, 3.0), random_flow(230, 310, 8, 3.0)
    f1[:100, :100] = 0
    r.tile_size, r.overlap = 100, 20
    merged = r._merge_flow_in_tiles(f1.copy(), f2.copy())
    np.savez_compressed(os.path.join(OUT, "merge.npz"), f1=f1, f2=f2, merged=merged, T=100, ov=20, **v)

    # 6. end to end register() + warp() with stdout
    ref, mov = synth_pair(420, 500, 9, np.uint16)
    r = mod.OptFlowRegistrator()
    r.num_pyr_lvl, r.num_iterations, r.tile_size, r.overlap, r.use_full_res_img, r.use_dog = 1, 2, 150, 20, True, True
    r.ref_img, r.mov_img = ref, mov
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        flow = r.register()
    w = mod.Warper()
    w.tile_size, w.overlap = 150, 20
    w.image, w.flow = mov, flow
    warped = w.warp()
    np.savez_compressed(os.path.join(OUT, "e2e_small.npz"), ref=ref, mov=mov, flow=flow.astype(np.float32), warped=warped,
                        stdout=buf.getvalue(), num_pyr_lvl=1, num_iterations=2, tile_size=150, overlap=20,
                        use_full_res_img=True, use_dog=True, **v)
    # 7. the rarely taken 'Worse alignment' branches, forced in the unmodified reference (gate monkeypatched)
    ref, mov = synth_pair(360, 440, 10, np.uint16)
    forced = {}
    for name, dec, full_res in (("tft_full", (True, False, True), True), ("ft_nofull", (False, True), False), ("tf_nofull", (True, False), False)):
        it = iter(dec)
        old = ofr.check_if_higher_similarity
        ofr.check_if_higher_similarity = lambda *a, **k: [next(it)]
        try:
            r = mod.OptFlowRegistrator()
            r.num_pyr_lvl, r.num_iterations, r.tile_size, r.overlap, r.use_full_res_img = 2, 1, 120, 16, full_res
            r.ref_img, r.mov_img = ref, mov
            with contextlib.redirect_stdout(io.StringIO()):
                forced[name] = r.register()
        finally:
            ofr.check_if_higher_similarity = old
    np.savez_compressed(os.path.join(OUT, "forced_decisions.npz"), ref=ref, mov=mov, num_pyr_lvl=2, num_iterations=1,
                        tile_size=120, overlap=16, **forced, **v)
    pipeline_goldens(v)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
