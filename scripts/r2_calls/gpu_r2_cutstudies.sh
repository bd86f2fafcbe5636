# all cut-cell studies of the reference on the GPU (IP and local DG; cylinder, plane interface, curved interface)
set -x
mkdir -p gpurun_out
timeout 900 python examples/cut_interface_convergence.py > gpurun_out/r2_cut_interface_studies_gpu.log 2>&1
cat gpurun_out/r2_cut_interface_studies_gpu.log
timeout 900 python -m pytest tests/test_cutcell_gpu.py -q -x 2>&1 | tail -n 12
