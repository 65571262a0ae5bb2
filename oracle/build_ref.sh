#!/bin/sh
# oracle/build_ref.sh <reference-root>
#
# Compiles the UNMODIFIED reference sources, where they lie, into oracle/_ref/
# (git-ignored; travels to the GPU box with the snapshot).  Nothing is copied
# out of the reference tree.  Each <test>/<test>.c holds kernel + main(); we
# rename both with -D so all tests and both precisions link into one library:
#     kernel <test>  ->  kgref_<test>_<f|d>
#     main           ->  kgref_main_<test>_<f|d>   (kept, unused)
# jacobi's kernel is Fortran (jacobi/jacobi.F90; no gfortran here): only its C
# driver (jacobi/main.c) is built, linked against the oracle's restatement
# through ref_shims.c.
set -e
REF="$1"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT/obj" "$OUT/bin"

CTESTS="laplacian wave13pt divergence gradient uxx1 lapgsrb gaussblur gameoflife tricubic tricubic2 vecadd matvec"
BASE="-D_GNU_SOURCE -std=c99 -fPIC -w"

build_lib() {  # name, flags
    name="$1"; shift
    flags="$*"
    objs=""
    for t in $CTESTS; do
        for r in float double; do
            s=$( [ $r = float ] && echo f || echo d )
            o="$OUT/obj/${name}_${t}_${s}.o"
            gcc $flags $BASE -Dreal=$r -D$t=kgref_${t}_${s} -Dmain=kgref_main_${t}_${s} \
                -I"$REF/$t" -c "$REF/$t/$t.c" -o "$o"
            objs="$objs $o"
        done
    done
    gcc $flags $BASE -c "$REF/laplacian/timing.c" -o "$OUT/obj/${name}_timing.o"
    gcc -shared $flags -o "$OUT/libkgref_${name}.so" $objs "$OUT/obj/${name}_timing.o" -lrt -lm
    echo "built $OUT/libkgref_${name}.so"
}

# exact shipped flags: makefile.in:28-29
build_lib shipped -O3 -ffast-math -march=native
build_lib strict  -O2 -fno-fast-math -ffp-contract=off
build_lib omp     -O3 -ffast-math -march=x86-64-v3 -fopenmp

# the reference gcc-target drivers themselves (text-level golden: i_mean/f_mean)
for t in $CTESTS; do
    for r in float double; do
        gcc -O3 -ffast-math -march=native -D_GNU_SOURCE -std=c99 -w -Dreal=$r \
            -I"$REF/$t" "$REF/$t/$t.c" "$REF/$t/timing.c" -o "$OUT/bin/${t}_${r}" -lrt -lm
    done
done
# jacobi: reference C driver + restated kernel (by-reference Fortran ABI, jacobi/main.c:43-52,245)
for r in float double; do
    s=$( [ $r = float ] && echo f || echo d )
    gcc -O3 -ffast-math -march=native -D_GNU_SOURCE -std=c99 -w -Dreal=$r \
        -I"$REF/jacobi" -I"$HERE" "$REF/jacobi/main.c" "$REF/jacobi/timing.c" \
        -DKGREF_SHIM_SFX=$s "$HERE/ref_shims.c" "$HERE/kgo.c" -o "$OUT/bin/jacobi_${r}" -lrt -lm
done

# the reference's own naive CUDA target (<test>/cuda/makefile:36-49), rebuilt for sm_100: the GPU-vs-GPU
# baseline ("the recompiled kernels") that tools/ref_cuda_table.py times beside the b200 kernels
if command -v nvcc >/dev/null 2>&1; then
    mkdir -p "$OUT/cuda_bin" "$OUT/obj_cuda"
    for t in $CTESTS; do
        [ -d "$REF/$t/cuda" ] || continue
        for r in float double; do
            o="$OUT/obj_cuda/${t}_${r}"
            ( nvcc -w -I"$REF/$t/cuda" -I"$REF/$t" -O3 -arch=sm_100 -Dreal=$r -c "$REF/$t/$t.cu" -o "$o.o" &&
              nvcc -w -I"$REF/$t/cuda" -I"$REF/$t" -O3 -arch=sm_100 -c "$REF/$t/cuda/cuda_profiling.cu" -o "$o.prof.o" &&
              nvcc -w -O3 -arch=sm_100 -x c -c "$REF/$t/timing.c" -o "$o.timing.o" &&
              nvcc -arch=sm_100 "$o.o" "$o.timing.o" "$o.prof.o" -o "$OUT/cuda_bin/${t}_${r}" -lrt \
                  -Xlinker --wrap=cudaLaunch -Xlinker --wrap=cudaLaunchKernel -Xcompiler -rdynamic --cudart=shared -ldl &&
              cp "$REF/$t/cuda/kernel" "$OUT/cuda_bin/${t}.kernel" ) 2>/dev/null || echo "reference cuda target of $t ($r) does not build"
        done
    done
    echo "built $OUT/cuda_bin"
fi

grep -m1 '^flags' /proc/cpuinfo | cut -d: -f2 > "$OUT/build_host_flags.txt"
rm -rf "$OUT/obj" "$OUT/obj_cuda"
echo "oracle/_ref complete"

# A writable checkout of the whole suite for the `benchmark`-on-the-GPU test (tests/test_gpu_benchmark.py): the Perl
# scripts, makefiles and test sources the real `./benchmark <nx> <ny> <ns> <nt> <nruns> b200 gcc` needs on the GPU box.
# Staged under oracle/_ref (git-ignored: never committed), it travels with the snapshot like the other _ref artefacts.
rm -rf "$OUT/suite"
mkdir -p "$OUT/suite"
( cd "$REF" && tar --exclude=.git -cf - . ) | ( cd "$OUT/suite" && tar -xf - )
chmod -R u+w "$OUT/suite"
echo "staged $OUT/suite"
