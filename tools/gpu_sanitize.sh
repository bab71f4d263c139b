#!/bin/bash
# compute-sanitizer passes over the kernel unit tests and one small engine forward (VERDICT r01 task 9a).
#   tools/gpu_sanitize.sh TAG
# memcheck: every unit test of the stand-alone kernels and of the tcgen05 building blocks + one CIFAR-10-sized forward in
# both GroupNorm plans.  racecheck (shared-memory hazards): the CUDA-core kernels only - the tool does not model
# mbarrier / TMA / tcgen05 ordering, so the tensor-core kernels are checked by memcheck and by the parity tests instead.
TAG=${1:-x}
mkdir -p gpurun_out; S=gpurun_out/sanitize_$TAG.txt; rm -f $S
run() {  # name tool timeout kexpr files...
  local name=$1 tool=$2 to=$3 k=$4; shift 4
  timeout $to compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest "$@" -q -m gpu -x --tb=line -p no:cacheprovider -k "$k" > gpurun_out/${name}_${TAG}.log 2>&1
  local rc=$?
  echo "$name ($tool, -k '$k'): exit $rc; $(grep -h 'passed\|failed' gpurun_out/${name}_${TAG}.log | tail -1); $(grep -h 'ERROR SUMMARY' gpurun_out/${name}_${TAG}.log | tail -1)" >> $S
}
run memcheck_kernels memcheck 400 "not randn" tests/test_gpu_kernels.py
run memcheck_tc memcheck 600 "not cifar10 and not full_1000 and not K_steps" tests/test_gpu_tc.py
run memcheck_engine memcheck 500 "groupnorm_on_load and 2" tests/test_gpu_round2.py
run racecheck_kernels racecheck 500 "groupnorm or upfirdn or fused or softmax or attn_small" tests/test_gpu_kernels.py
cat $S
