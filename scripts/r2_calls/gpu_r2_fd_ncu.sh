# ncu capture of the FD kernel with the one-iteration-ahead prefetch (the r1 capture under profiles/ is the old kernel)
set -x
mkdir -p gpurun_out
timeout 200 python scripts/quick_bench.py --fd 1 --reps 2 > gpurun_out/r2f_plain_fd.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fd_apply_vec -s 2 -c 2 -f -o gpurun_out/r2f_fd \
    python scripts/quick_bench.py --fd 1 --reps 2 > gpurun_out/r2f_ncu_fd.log 2>&1
tail -n 3 gpurun_out/r2f_plain_fd.log; ls -la gpurun_out/r2f_fd.ncu-rep
