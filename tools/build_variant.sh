#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D flags...] — build variants/libare_b200_NAME.so with extra defines for render.cu (A/B runs:
# ARE_B200_LIB=$PWD/variants/libare_b200_NAME.so python bench.py ...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
CS=aurora_rendering_engine_b200/csrc; OBJ=aurora_rendering_engine_b200/lib/obj
mkdir -p variants
make -C $CS -j4 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -prec-div=false -prec-sqrt=false -ftz=true "$@" -Xptxas -v -c $CS/render.cu -o variants/render_$name.o 2> variants/ptxas_$name.txt
grep -A2 "k_render_pathILi0ELb0ELb0" variants/ptxas_$name.txt | tail -2 | tr '\n' ' '; echo
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libare_b200_$name.so variants/render_$name.o $OBJ/wavefront.o $OBJ/harness64.o $OBJ/rtao.o $OBJ/api.o $OBJ/scene.o $OBJ/patch.o $OBJ/lbvh.o $OBJ/bake.o $OBJ/embed.o -lcudart_static -lpthread -ldl -lrt
