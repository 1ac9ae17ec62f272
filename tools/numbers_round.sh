#!/bin/bash
# Numbers for DESIGN.md section 6 / profiles: every bench workload, the per-kernel micro-benchmark, the PyTorch-eager
# context number.  usage: tools/numbers_round.sh <tag>
tag=${1:-r01}
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg2.log 2> gpurun_out/${tag}_bench_cfg2.err
cp gpurun_out/bench_kernels.json gpurun_out/${tag}_bench_kernels.json
tail -n 1 gpurun_out/${tag}_bench_cfg2.log | cut -c1-140
for wl in cfg1_7m_32_fp32 cfg3_14m_64_bf16 cfg4_22m_tin_64_bf16 cfg5_model_b_eval; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-profile --no-cpu-baseline > gpurun_out/${tag}_bench_$wl.log 2>&1
  echo "$wl: $(tail -n 1 gpurun_out/${tag}_bench_$wl.log | cut -c1-140)"
done
python tools/kbench.py --reps 20 --out gpurun_out/${tag}_kbench.json > gpurun_out/${tag}_kbench.log 2>&1
python tools/torch_eager_bench.py --steps 3 > gpurun_out/${tag}_eager.log 2>&1; tail -n 1 gpurun_out/${tag}_eager.log
