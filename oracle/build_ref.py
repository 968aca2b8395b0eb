"""Recipe: BUILD the unmodified reference (pure Python) into oracle/_ref/ so that it travels to the GPU box.

The reference has no native code; its "binary" is CPython bytecode.  This recipe byte-compiles the reference's own
source files where they lie under /root/reference (py_compile, sourceless layout, suffix .pyb) — the Python analogue of
compiling a C reference into oracle/_ref/*.so.  No reference source text is copied into the repo.

    python oracle/build_ref.py          # needs /root/reference (build container only); idempotent

oracle/_ref/ is git-ignored but NOT gpurun-ignored, exactly like a compiled oracle/_ref/*.so would be for a C reference.  Users: `bench.py --impl reference` (times the reference's own
UNetp on the host cores, cpu_baseline.kind "reference") and tests/test_reference_drivers_gpu.py (runs the reference's
own train.train / eval.eval_net / eval.score_model_best_iou / infer.inference against the drop-in `unet` package and
against the reference's `unet` package side by side).  The product path never imports anything from here.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref", "src")
FILES = ["unet/__init__.py", "unet/unet_p.py", "unet/unet_p_res.py", "train.py", "eval.py", "infer.py",
         "utils/__init__.py", "utils/data_set.py", "utils/img_utils.py", "utils/iou_metric.py", "utils/rle_encode.py",
         "utils/data_visualization.py", "utils/keras_history_visualization.py"]


def build(quiet=False):
    if not os.path.isdir(SRC):
        if not quiet:
            print("oracle/build_ref.py: %s not present (GPU box?) - keeping whatever oracle/_ref holds" % SRC)
        return os.path.isdir(DST)
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            continue
        # X.py -> X.pyb (CPython bytecode, same format as a .pyc; the gpurun snapshot drops *.pyc) next to where the source
        # would be; oracle/ref_loader.py registers a sourceless loader for that suffix
        dst = os.path.join(DST, rel[:-3] + ".pyb")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile="reference/src/" + rel, doraise=True)
    with open(os.path.join(HERE, "_ref", "README"), "w") as f:
        f.write("Bytecode of yaricom/Plastic-UNet src/ built by oracle/build_ref.py with %s; git-ignored test infrastructure.\n"
                % sys.version.split()[0])
    return True


def available():
    return os.path.exists(os.path.join(DST, "unet", "unet_p.pyb"))


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref staged" if ok else "oracle/_ref unavailable")
    sys.exit(0)
