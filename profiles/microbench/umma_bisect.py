import os, sys, torch
sys.path.insert(0, '/root/repo')
import amt_saga_b200
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(600), 264600, seed_base=50000)
plan = ops.get_cqt_plan(44100, 512, note_to_hz('C1'), 84, 12, 2)
def timeit(n=10):
    for _ in range(3): ops.cqt_batch(wav, plan, impl=2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): ops.cqt_batch(wav, plan, impl=2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
    os.environ['SAGA_UMMA_DEBUG'] = str(dbg)
    print('debug mask %d (1=noMMA 2=noProducerData 4=noEpilogue): cqt stage %.3f ms (cascade ~0.74 ms included)' % (dbg, timeit()))
