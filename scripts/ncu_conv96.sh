python scripts/conv_diag.py floor2 2 96 96 > gpurun_out/plain_c96.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_conv_umma -s 1 -c 1 -o gpurun_out/r02_conv96_c python scripts/conv_diag.py floor2 2 96 96 > gpurun_out/ncu_c96_c.log 2>&1
MK3D_DEBUG=7 python scripts/conv_diag.py floor2 2 96 96 > gpurun_out/plain_c96_d7.log 2>&1 && \
MK3D_DEBUG=7 ncu --set full --clock-control none --import-source on -k regex:k_conv_umma -s 1 -c 1 -o gpurun_out/r02_conv96_c_d7 python scripts/conv_diag.py floor2 2 96 96 > gpurun_out/ncu_c96_c_d7.log 2>&1
tail -3 gpurun_out/ncu_c96_c.log gpurun_out/ncu_c96_c_d7.log
