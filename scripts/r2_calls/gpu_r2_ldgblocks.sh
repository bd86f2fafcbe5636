# local-DG element-block path: its GPU tests, the cut-cell LDG cylinder study, and the tests of everything that
# was refactored around it (Schur solver header, COO kernel, CG on the Krylov toolkit, p = 2 fused step defaults)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cutcell_gpu.py -q -x -k "ldg" 2>&1 | tail -n 15 > gpurun_out/r2_ldgblocks_tests.log
tail -n 4 gpurun_out/r2_ldgblocks_tests.log
timeout 600 python examples/cylinder_ldg_convergence.py > gpurun_out/r2_ldg_cylinder_study_gpu.log 2>&1
cat gpurun_out/r2_ldg_cylinder_study_gpu.log
timeout 900 python -m pytest tests/test_cutcell_gpu.py tests/test_ldg_gpu.py tests/test_ipdg_gpu.py -q -x -k "not ldg_cut and not ldg_cylinder and not ldg_blocks" 2>&1 | tail -n 4
