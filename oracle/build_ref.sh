#!/usr/bin/env bash
# Build the UNMODIFIED reference rasterizer (submodules/diff-gaussian-rasterization) and
# simple-knn straight from the sources where they lie under /root/reference, for sm_100,
# into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
#
# This is TEST/BENCH INFRASTRUCTURE: it is the pin for the parity tests and the
# "stock CUDA rasterizer" arm of bench.py. Nothing in street_gaussians_b200/ loads it.
#
# No reference source is copied into the repo history. gcc 13 no longer leaks <cstdint>/<cfloat>
# through other headers, so the two missing includes are injected from the command line
# (-include cstdint / -include cfloat) instead of patching the sources.
# The package __init__.py files are installed next to the built _C.so exactly as
# `pip install --target` would do (install output, git-ignored).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
DGR="$REF/submodules/diff-gaussian-rasterization"
KNN="$REF/submodules/simple-knn"
if [ ! -d "$DGR" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; fi

PY=${PYTHON:-python}
TORCH_INC=$($PY - <<'PY'
import torch.utils.cpp_extension as c, sysconfig, warnings
print(" ".join("-I"+p for p in c.include_paths() + [sysconfig.get_paths()["include"]]))
PY
)
TORCH_LIB=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))" 2>/dev/null)
EXT_SUFFIX=".so"
COMMON="-O3 -std=c++17 -D_GLIBCXX_USE_CXX11_ABI=1 -DTORCH_API_INCLUDE_EXTENSION_H -DTORCH_EXTENSION_NAME=_C $TORCH_INC"
NVCCF="-gencode arch=compute_100,code=sm_100 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fno-gnu-unique \
 -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_BFLOAT16_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__ -w"
LINK="-shared -L$TORCH_LIB -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda -L/usr/local/cuda/lib64 -lcudart -Xlinker -rpath -Xlinker $TORCH_LIB"

build_dgr() {
  local pkg="$OUT/ref_dgr" obj="$OUT/obj_dgr"
  mkdir -p "$pkg" "$obj"
  local pids=()
  for f in cuda_rasterizer/rasterizer_impl.cu cuda_rasterizer/forward.cu cuda_rasterizer/backward.cu rasterize_points.cu; do
    nvcc -c "$DGR/$f" -o "$obj/$(basename $f).o" $COMMON $NVCCF -include cstdint -I"$DGR/third_party/glm" -I"$DGR" &
    pids+=($!)
  done
  g++ -c "$DGR/ext.cpp" -o "$obj/ext.o" $COMMON -fPIC -I/usr/local/cuda/include -I"$DGR" -w &
  pids+=($!)
  for p in "${pids[@]}"; do wait $p; done
  nvcc $obj/*.o -o "$pkg/_C$EXT_SUFFIX" $LINK
  install -m 644 "$DGR/diff_gaussian_rasterization/__init__.py" "$pkg/__init__.py"
}
build_knn() {
  local pkg="$OUT/ref_knn" obj="$OUT/obj_knn"
  mkdir -p "$pkg" "$obj"
  local pids=()
  for f in simple_knn.cu spatial.cu; do
    nvcc -c "$KNN/$f" -o "$obj/$f.o" $COMMON $NVCCF -include cfloat -I"$KNN" &
    pids+=($!)
  done
  g++ -c "$KNN/ext.cpp" -o "$obj/ext.o" $COMMON -fPIC -I/usr/local/cuda/include -I"$KNN" -w &
  pids+=($!)
  for p in "${pids[@]}"; do wait $p; done
  nvcc $obj/*.o -o "$pkg/_C$EXT_SUFFIX" $LINK
  : > "$pkg/__init__.py"
}
build_dgr
build_knn
rm -rf "$OUT/obj_dgr" "$OUT/obj_knn"
echo "built: $(ls $OUT/*/_C.so)"
