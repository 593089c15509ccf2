"""Stage timings of the progressive path on 4K 4:2:0 q75 SOF2 frames (BASELINE.json configs[3] geometry). The frames are
made in the build container by the reference encoder (`oracle/_ref/jpeg -q 75 -v -s 1x1,2x2,2x2 -z 240`) into
libjpeg_b200/build/prog4k/ (not tracked)."""
import glob, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import libjpeg_b200
base = [open(p, 'rb').read() for p in sorted(glob.glob('libjpeg_b200/build/prog4k/*.jpg'))]
if not base:
    sys.exit('no frames under libjpeg_b200/build/prog4k/: make them in the build container first (see the docstring)')
print('frames', [len(b) for b in base])
for nf in [int(a) for a in sys.argv[1:]] or [64, 256]:
    frames = [base[i % len(base)] for i in range(nf)]
    dec = libjpeg_b200.BatchDecoder(frames)
    out = dec.new_output(); dec.upload(); dec.enable_timing(True)
    for it in range(3):
        dec.decode(out); torch.cuda.synchronize()
        e, r = dec.last_timing()
        print(nf, 'frames: entropy %.3f ms recon %.3f ms -> %.0f fps, launches %d' % (e, r, nf / (e + r) * 1e3, dec.launches), flush=True)
    print('status', [dec.status(i) for i in range(min(nf, 4))])
    del dec, out
    torch.cuda.empty_cache()
