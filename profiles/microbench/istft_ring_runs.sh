for r in 48 64 87 96 130 174 260; do echo "run_cap $r: $(SAGA_ISTFT_RING_RUN=$r python profiles/microbench/istft_k4_only.py)"; done
