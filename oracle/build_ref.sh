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

grep -m1 '^flags' /proc/cpuinfo | cut -d: -f2 > "$OUT/build_host_flags.txt"
rm -rf "$OUT/obj"
echo "oracle/_ref complete"
