"""Stage timings (CUDA events recorded by the library around the entropy and the reconstruction stage) of one workload at the
given batch sizes:  python tools/probe.py [--workload cfg3] [sizes ...]   (B200JPG_LIB selects a kernel variant build)"""
import sys, time
sys.path.insert(0, '.')
import torch
import libjpeg_b200
from tools import bench_inputs
args = sys.argv[1:]
workload = 'cfg3'
if args and args[0] == '--workload':
    workload = args[1]
    args = args[2:]
base = bench_inputs.make_frames(workload, 4, 4)
print(workload, 'lib', libjpeg_b200.library_path(), 'bytes', [len(b) for b in base], flush=True)
for nf in [int(a) for a in args] or [128, 512]:
    frames = [base[i % 4] for i in range(nf)]
    dec = libjpeg_b200.BatchDecoder(frames)
    out = dec.new_output(); dec.upload(); dec.enable_timing(True)
    best = None
    for it in range(5):
        dec.decode(out); torch.cuda.synchronize()
        e, r = dec.last_timing()
        best = (e, r) if best is None or e + r < sum(best) else best
    e, r = best
    print('%s %d frames: entropy %.3f ms recon %.3f ms -> %.0f fps (entropy only %.0f, recon only %.0f), launches %d' % (workload, nf, e, r, nf / (e + r) * 1e3, nf / e * 1e3, nf / r * 1e3, dec.launches), flush=True)
    assert all(dec.status(i) == 0 for i in range(min(nf, 8)))
    del dec, out
    torch.cuda.empty_cache()
