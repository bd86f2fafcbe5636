# phase-aware tile order (STEFAN_TILE_PAIRING): co-resident CTAs on the same phase body -- A/B on the two-phase 100 M-dof meshes
set -x
mkdir -p gpurun_out
for pr in 0 1 2; do
  echo "=== STEFAN_TILE_PAIRING=$pr"
  for o in 2 3 1; do STEFAN_TILE_PAIRING=$pr timeout 120 python scripts/quick_bench.py --order $o --reps 30; done
done > gpurun_out/r2_tile_pairing.log 2>&1
grep -v "^+" gpurun_out/r2_tile_pairing.log
