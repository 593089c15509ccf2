import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import libjpeg_b200
from libjpeg_b200 import synth
t0=time.time()
base=[synth.frame(3840,2160,s) for s in range(1,5)]
print('gen 4 frames', time.time()-t0, 'bytes', [len(b) for b in base], flush=True)
sizes = [int(a) for a in sys.argv[1:]] or [128, 512]
for nf in sizes:
    frames=[base[i%4] for i in range(nf)]
    t0=time.time(); dec=libjpeg_b200.BatchDecoder(frames); print('batch_create', nf, time.time()-t0, flush=True)
    out=dec.new_output(); dec.upload(); dec.enable_timing(True)
    for it in range(4):
        dec.decode(out); torch.cuda.synchronize()
        e,r=dec.last_timing()
        print(nf,'frames: entropy %.3f ms recon %.3f ms -> %.0f fps (entropy only %.0f, recon only %.0f)'%(e,r,nf/(e+r)*1e3, nf/e*1e3, nf/r*1e3), flush=True)
    print('status', [dec.status(i) for i in range(min(nf,4))])
    algo_a = dec.ecs_bytes + 128*dec.stored_blocks
    print('entropy algorithmic GB/s', algo_a/ (e*1e-3)/1e9, 'of 6485')
    del dec, out
    torch.cuda.empty_cache()
