#!/bin/bash
# e2e pipeline sweep on the GPU box: first-chunk size (zero-copy outputs) and the staged variants, with per-chunk traces
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-postorder --no-cpu"
for cfg in "65536" "32768" "131072" "16384"; do
  echo "== zero-copy, chunk $cfg" >> gpurun_out/e2e_sweep.log
  PCGDO_HOST_CHUNK_PAIRS=$cfg $B 2>/dev/null | grep -o '"value.*GCUPS", "n_gpus\|"e2e.*gpu_launches' >> gpurun_out/e2e_sweep.log
done
echo "== trace zero-copy chunk 65536" >> gpurun_out/e2e_sweep.log
PCGDO_TRACE_PIPELINE=1 $B 2>&1 | grep -A20 "pcgdo pipeline" | tail -14 >> gpurun_out/e2e_sweep.log
echo "== staged compact" >> gpurun_out/e2e_sweep.log
PCGDO_NO_ZERO_COPY=1 $B 2>/dev/null | grep -o '"e2e.*gpu_launches' >> gpurun_out/e2e_sweep.log
echo "== no compact" >> gpurun_out/e2e_sweep.log
$B --no-compact 2>/dev/null | grep -o '"e2e.*gpu_launches' >> gpurun_out/e2e_sweep.log
