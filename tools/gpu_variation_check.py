"""one-shot GPU check of off-default parameters (patch splitting etc.): product vs oracle, prints mismatches"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bindings, synth
t0 = time.time()
frames = [synth.figure(scale=0.12, seed=3, frame=0), synth.double_sheet(n_side=32, seed=5)]
prod, orc = bindings.Product(0), bindings.Oracle()
w = prod.weight_normal(frames[0][0], 11)
for name, prec, kw in (("split64", 4, dict(max_patch_size=64)), ("prec1_thin2_lvl32_cc8", 1, dict(surface_thickness=2, min_level=32, min_point_count_per_cc=8)),
                       ("lambda1_r96_noorient_nosplit", 4, dict(lambda_refine=1.0, search_radius_refine=96, normal_orientation=0, enable_patch_splitting=0))):
    prm = bindings.ctc_seg_params(bits=10, iterations=3, weight=w)
    for k, v in kw.items():
        setattr(prm, k, v)
    try:
        bad = bindings.compare_gof(prod.encode_gof(frames, prm, occupancy_precision=prec), orc.encode_gof(frames, prm, occupancy_precision=prec))
        print(name, "OK" if not bad else "MISMATCH", bad[:5], "t=%.1f" % (time.time() - t0), flush=True)
    except Exception as e:
        print(name, "EXC", e, flush=True)
