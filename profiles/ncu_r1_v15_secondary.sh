# ncu --set full captures of the kernels behind the reference-default / per-note shapes (final build):
# stft_eo4096_kernel (n_fft 4096 forward), istft_kernel (2048 and 4096), cqt_contract2_kernel (348 bins, 48 per octave).
set -x
python profiles/microbench/refdefault_pipeline.py > gpurun_out/plain15s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:stft_eo4096' -s 10 -c 2 -f -o gpurun_out/prof_v15_eo python profiles/microbench/refdefault_pipeline.py > gpurun_out/ncu15a.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:cqt_contract2' -s 8 -c 1 -f -o gpurun_out/prof_v15_ct2 python profiles/microbench/cqt_contract2_probe.py > gpurun_out/ncu15b.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:istft_kernel' -s 2 -c 1 -f -o gpurun_out/prof_v15_istft python profiles/microbench/istft_probe.py > gpurun_out/ncu15c.log 2>&1
ls -la gpurun_out/*.ncu-rep
