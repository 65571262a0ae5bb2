#!/bin/sh
# install_overlay.sh <suite-root> [<b200-repo-root>]
#
# Adds the `b200` target to a WRITABLE checkout of kernelgen-perf-tests, next to gcc/cuda:
#   * <test>/b200/{makefile,kernel} for every test the library implements
#     (benchmark:311-333 discovers a test by the existence of <test>/<target>/),
#   * `b200`, `%.b200`, `%.b200.clean`, `test.b200` rules appended to the top makefile
#     (same shape as the %.cuda rules, makefile:74-120),
#   * `benchmark` taught that b200 is a cuda-like target, so it reads <test>/b200/kernel, passes
#     PROFILING_FNAME and fills the t_krn / nreg_krn columns (benchmark:126,162,174,210).
# Nothing of the suite is copied into this repository; the patch is applied to the user's tree.
set -e
SUITE="$1"
HERE="$(cd "$(dirname "$0")" && pwd)"
B200_ROOT="${2:-$(cd "$HERE/../.." && pwd)}"
[ -f "$SUITE/benchmark" ] && [ -f "$SUITE/makefile.in" ] || { echo "usage: $0 <kernelgen-perf-tests checkout>"; exit 1; }

TESTS3D="laplacian wave13pt divergence gradient uxx1 lapgsrb tricubic tricubic2 vecadd sincos matmul"
TESTS2D="jacobi gaussblur gameoflife matvec"

for t in $TESTS3D $TESTS2D; do
    [ -d "$SUITE/$t" ] || continue
    d="$SUITE/$t/b200"
    mkdir -p "$d"
    printf '%s' "$t" > "$d/kernel"
    T=$(echo "$t" | tr a-z A-Z)
    case " $TESTS2D " in *" $t "*) RUNARGS='$(NX) $$(($(NY) * $(NS))) $(NT)';; *) RUNARGS='$(NX) $(NY) $(NS) $(NT)';; esac
    cat > "$d/makefile" <<MK
##  b200 target of the $t test: the host driver is kernelgen-perf-tests_b200/drivers/driver.c,
##  the kernels live in libb200stencil.so (hand-written sm_100a CUDA, no CPU fallback).
-include ../../makefile.in

VERBOSE ?= 0
SILENT ?= @
ifeq (\$(VERBOSE), 1)
SILENT =
endif

real ?= float
GCC ?= gcc -O3 -D_GNU_SOURCE -std=c99
B200_ROOT ?= $B200_ROOT
B200_PKG = \$(B200_ROOT)/kernelgen-perf-tests_b200
NX ?= 512
NY ?= 256
NS ?= 256
NT ?= 10

all: $t

\$(B200_PKG)/libb200stencil.so:
	\$(SILENT)\$(MAKE) -C \$(B200_PKG)/csrc

B200_SRCS = \$(B200_PKG)/drivers/driver.c \$(B200_PKG)/drivers/timing.c \$(B200_PKG)/drivers/kg_rand.c

$t: \$(B200_SRCS) \$(B200_PKG)/drivers/kg_init.h \$(B200_PKG)/libb200stencil.so
	\$(SILENT)\$(GCC) -Dreal=\$(real) -DB200_TEST_ID=B200_$T -I\$(B200_ROOT)/include -I\$(B200_PKG)/drivers \\
		\$(B200_SRCS) -o \$@ \\
		-L\$(B200_PKG) -lb200stencil -Wl,-rpath,\$(B200_PKG) -lrt -lm -lpthread

clean:
	\$(SILENT)rm -rf *.o $t

# the measured HBM ceiling of this pool (driver-written MEASURED_PEAKS.json), for the "% of the HBM roofline" figure
B200_HBM_GBS ?= \$(shell sed -n 's/.*"hbm_gbs": *\([0-9.]*\).*/\1/p' \$(B200_ROOT)/MEASURED_PEAKS.json 2>/dev/null)
export B200_HBM_GBS

test: $t
	\$(SILENT)./\$< $RUNARGS

# like <test>/cuda/makefile:54-66 of the suite: check = memory checker, roofline = profiler metrics of the kernel
SMALLARGS = 64 32 $( [ "$RUNARGS" = '$(NX) $(NY) $(NS) $(NT)' ] && echo 32 ) 2
check: $t
	\$(SILENT)compute-sanitizer --tool memcheck --error-exitcode 1 ./\$< \$(SMALLARGS)

racecheck: $t
	\$(SILENT)compute-sanitizer --tool racecheck --error-exitcode 1 ./\$< \$(SMALLARGS)

synccheck: $t
	\$(SILENT)compute-sanitizer --tool synccheck --error-exitcode 1 ./\$< \$(SMALLARGS)

roofline: $t
	\$(SILENT)ncu --clock-control none -c 1 -s 2 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum ./\$< $RUNARGS
MK
done

# top-level makefile: append the b200 fan-out rules once
if ! grep -q 'TARGETS_B200' "$SUITE/makefile"; then
    cat >> "$SUITE/makefile" <<'MK'

## ---- b200 target (added by kernelgen-perf-tests_b200/suite_overlay/install_overlay.sh) ----
B200_TESTS = $(TARGETS) jacobi sincos matmul
TARGETS_B200 = $(addsuffix .b200, $(B200_TESTS))
TARGETS_B200_CLEAN = $(addsuffix .b200.clean, $(B200_TESTS))

.PHONY: b200 b200.clean test.b200

b200: $(TARGETS_B200)

b200.clean: $(TARGETS_B200_CLEAN)

%.b200:
	$(SILENT)if [ -d $(subst .b200,,$@)/b200 ]; then cd $(subst .b200,,$@)/b200 && $(MAKE); fi

%.b200.clean:
	$(SILENT)if [ -d $(subst .b200.clean,,$@)/b200 ]; then cd $(subst .b200.clean,,$@)/b200 && $(MAKE) clean; fi

test.b200: $(TARGETS_B200)
	$(SILENT)./benchmark $(NX) $(NY) $(NS) $(NT) $(NRUNS) b200
MK
fi

# benchmark: treat b200 like the cuda targets (kernel file, PROFILING_FNAME, t_krn, nreg_krn) ...
if ! grep -q 'b200' "$SUITE/benchmark"; then
    sed -i 's/(\$target =~ m\/\^cuda\.\*\$\/)/(($target =~ m\/^cuda.*$\/) or ($target eq "b200"))/g' "$SUITE/benchmark"
    # ... and append two columns to its table, GLUP/s and % of the HBM roofline, parsed from the line the
    # b200 driver prints ("b200: 1 GPU(s), X GLUP/s, Y GB/s algorithmic ..., Z % of the HBM roofline").  Other targets: N/A.
    # mktable / mkchart read the "Data for table" / "Times for chart" blocks, which keep their format.
    perl -0pi -e '
        s/(\tprint_field\(\$final_mean\);\n)/$1\tprint_field(find_average(\$output, "", qr{b200:\\s\\d+\\sGPU\\(s\\),\\s([-+]?[0-9]*\\.?[0-9]+([eE][-+]?[0-9]+)?)\\sGLUP\/s}));\n\tprint_field(find_average(\$output, "", qr{,\\s([-+]?[0-9]*\\.?[0-9]+([eE][-+]?[0-9]+)?)\\s%\\sof\\sthe\\sHBM\\sroofline}));\n/;
        s/\| %8s \|\\n",\n(\t\t"test", "target", "i_mean", "t_init", "t_alloc", "t_load", "t_comp",\n\t\t"t_krn", "nreg_krn", "t_save", "t_free", "f_mean")\);/| %8s | %8s | %8s |\\n",\n$1, "GLUP\/s", "%roof");/g;
        s/my\(\$table_width\) = 148;/my(\$table_width) = 170;/;
    ' "$SUITE/benchmark"
fi
echo "b200 target installed into $SUITE (library root: $B200_ROOT)"
