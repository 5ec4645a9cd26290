#!/bin/bash
# e2e pipeline sweep on the GPU box: chunk size of the staged compact transfer, the zero-copy experiment, no packing,
# and a per-chunk trace of the default configuration (see DESIGN.md "Host entry point").
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-postorder --no-cpu"
for cfg in "65536" "32768" "131072" "16384"; do
  echo "== staged compact, chunk $cfg" >> gpurun_out/e2e_sweep.log
  PCGDO_HOST_CHUNK_PAIRS=$cfg $B 2>/dev/null | grep -o '"e2e.*gpu_launches' >> gpurun_out/e2e_sweep.log
done
echo "== trace (default)" >> gpurun_out/e2e_sweep.log
PCGDO_TRACE_PIPELINE=1 $B 2>&1 | grep -A22 "pcgdo pipeline" | tail -22 >> gpurun_out/e2e_sweep.log
echo "== zero-copy emit stores (experiment)" >> gpurun_out/e2e_sweep.log
PCGDO_ZERO_COPY=1 $B 2>/dev/null | grep -o '"e2e.*gpu_launches' >> gpurun_out/e2e_sweep.log
echo "== no compact" >> gpurun_out/e2e_sweep.log
$B --no-compact 2>/dev/null | grep -o '"e2e.*gpu_launches' >> gpurun_out/e2e_sweep.log
