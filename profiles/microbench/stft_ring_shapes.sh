for sh in 15 1520 19 1923; do
  echo "== shape $sh"; SAGA_STFT_RING=1 SAGA_STFT_RING_SHAPE=$sh timeout 200 python profiles/microbench/stft_ring_check.py 2>&1 | grep -v "^  \|^Traceback\|^Search\|^CUDA kernel\|^For debug\|^Compile with" | tail -7
done
