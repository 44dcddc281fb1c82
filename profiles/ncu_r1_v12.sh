# ncu evidence for the v12 build: launch list (all hot kernels, 3 warm-up + 2 timed steps) and full captures
set -x
B="python bench.py --steps 2 --warmup 3 --cpu-windows 0 --e2e-steps 3"
$B > gpurun_out/plain12.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:stft_kernel|cqt_|decimate|subtract_chain|window_db|pcm16' -c 120 --csv --log-file gpurun_out/launches_v12.csv $B > gpurun_out/ncu12a.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:stft_kernel|cqt_|decimate|subtract_chain|window_db' -s 39 -c 13 -f -o gpurun_out/prof_v12 $B > gpurun_out/ncu12b.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:pcm16' -s 8 -c 4 -f -o gpurun_out/prof_v12_pcm $B > gpurun_out/ncu12c.log 2>&1
tail -3 gpurun_out/ncu12a.log gpurun_out/ncu12b.log gpurun_out/ncu12c.log
ls -la gpurun_out/
