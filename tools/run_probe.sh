cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvcc -arch=sm_100a -O3 -o /tmp/lds_probe tools/lds_probe.cu 2>/dev/null
timeout 120 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum --csv --log-file gpurun_out/r02_lds_probe_ncu.csv /tmp/lds_probe 16 quick > gpurun_out/r02_lds_probe_names.txt 2>&1
tail -3 gpurun_out/r02_lds_probe_names.txt
