#!/bin/bash
# Build and run the REFERENCE tsadvc on the golden cases (needs gfortran; the graft image has none).
#   HYCOM_SRC=/path/to/HYCOM-src bash fortran/build_ref.sh [case ...]
# Produces tests/golden/from_reference.json, which tests/test_oracle.py::test_reference_fixture consumes:
# the oracle must reproduce the reference's bits.  (-ffp-contract=off: without it gfortran contracts
# a*b+c into FMAs under -march=native and the binary is only defined to ~1 ulp per operation)
# Also compiles the drop-in shim fortran/mod_tsadvc_b200.F90 against the reference's modules (syntax check
# of the binding; linking it needs libhycom_tsadvc_b200.so).
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(dirname "$HERE")
SRC=${HYCOM_SRC:-/root/reference}
FC=${FC:-gfortran}
OUT=$HERE/_ref; mkdir -p $OUT
EOSF=${EOS_FLAGS:--DEOS_SIG2 -DEOS_17T}          # sigver 6, the GLB builds (config/*relo*)
FFLAGS="-cpp -DREAL8 -DRELO -DTIMER -DENDIAN_IO $EOSF $EXTRA_DEFS -fdefault-real-8 -fdefault-double-8 -O2 -ffp-contract=off -fno-second-underscore -ffree-line-length-none -I$SRC -J$OUT"
cd $OUT
for f in mod_dimensions mod_xc; do $FC $FFLAGS -c $SRC/$f.F90 -o $f.o; done
$FC $FFLAGS -c $SRC/wtime.F90 -o wtime.o
gcc -O2 -c $SRC/machi_c.c -o machi_c.o 2>/dev/null || true
$FC $FFLAGS -c $SRC/mod_cb_arrays.F90 -o mod_cb_arrays.o
$FC $FFLAGS -c $HERE/mod_pipe_stub.F90 -o mod_pipe.o
$FC $FFLAGS -c $SRC/mod_tsadvc.F90 -o mod_tsadvc.o
$FC $FFLAGS -c $SRC/bigrid.F90 -o bigrid.o
$FC $FFLAGS -c $HERE/ref_driver.F90 -o ref_driver.o
$FC $FFLAGS -o ref_driver ref_driver.o mod_tsadvc.o bigrid.o mod_pipe.o mod_cb_arrays.o mod_xc.o mod_dimensions.o wtime.o $(ls machi_c.o 2>/dev/null)
# the shim: compile only (module name clashes with the reference's, so in its own directory)
mkdir -p $OUT/shim && (cd $OUT/shim && cp ../mod_dimensions.mod ../mod_xc.mod ../mod_cb_arrays.mod . && \
  $FC $FFLAGS -J$OUT/shim -c $HERE/mod_tsadvc_b200.F90 -o mod_tsadvc_b200.o && echo "shim compiles")
cd $ROOT
CASES=${@:-box_fct2 periodic_mpdata_tracers fct4_periodic_i pcm fct2c_btrmas isopyc diffusion_17t}
for c in $CASES; do
  python fortran/ref_case.py write $c $OUT/case_$c
  $OUT/ref_driver $OUT/case_$c
done
python fortran/ref_case.py digest $OUT $CASES
echo "wrote tests/golden/from_reference.json"
