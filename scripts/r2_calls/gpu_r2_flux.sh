# the LDG interface-flux driver on the GPU (n = 17; the last seconds of the round's GPU budget)
mkdir -p gpurun_out
timeout 50 python examples/cylinder_ldg_interface_flux.py 17 > gpurun_out/r2_ldg_interface_flux_gpu.log 2>&1
cat gpurun_out/r2_ldg_interface_flux_gpu.log
