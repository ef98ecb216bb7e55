#!/bin/sh
# make_ref.sh : stage the UNMODIFIED reference package into oracle/_ref/ (git-ignored, NOT gpurun-ignored,
# so it travels to the GPU box with the snapshot like the other built artefacts).
#
# TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is a copy of
# /root/reference/gpy_dla_detection/*.py (the path's modules; plottings/ and the CIV models are not on the path).
# bench.py --impl reference and the cpu_baseline leg run these files through oracle/ref_loader.py so that the CPU
# arm is the reference's own NullGP / SubDLAGP / DLAGP / BayesModelSelect / ZGP classes,
# cpu_baseline.kind = "reference".  Nothing under gpy_dla_detection_b200/ imports it.
# voigt.c (MEX + libcerf) cannot be compiled in this image (no mex.h, no cerf.h) and is not staged.
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/gpy_dla_detection" ]; then
  echo "make_ref.sh: $SRC/gpy_dla_detection not found (GPU box: the prebuilt oracle/_ref is used)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/gpy_dla_detection"
for f in __init__ bayesian_model_selection dla_gp dla_samples effective_optical_depth log_posterior_mcmc \
         model_priors null_gp set_parameters subdla_gp subdla_samples voigt voigt_lls zqso_gp zqso_samples \
         zqso_set_parameters; do
  cp "$SRC/gpy_dla_detection/$f.py" "$DST/gpy_dla_detection/$f.py"
done
( cd "$SRC/gpy_dla_detection" && sha256sum __init__.py bayesian_model_selection.py dla_gp.py null_gp.py subdla_gp.py \
    voigt.py zqso_gp.py ) > "$DST/SHA256SUMS"
echo "make_ref.sh: staged $(ls "$DST/gpy_dla_detection" | wc -l) files into $DST"
