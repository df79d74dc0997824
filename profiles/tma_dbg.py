import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import nvtt_b200_loader, refapi as ref
nvtt = nvtt_b200_loader.load()
ctx = nvtt.Context(0)
rng = np.random.default_rng(5)
w, h = 512, 256
im = rng.random((h, w, 4), dtype=np.float32)
for filt in (1, 2):
    a = ref.Surface(wrap=0); b = nvtt.Surface(ctx, wrap=0)
    a.set_image(2, w, h, im); b.set_image(2, w, h, im)
    a.build_next_mipmap(filt); b.build_next_mipmap(filt)
    ga, gb = a.get(), b.get()
    print(filt, ga.shape, int((ga.view(np.uint32) != gb.view(np.uint32)).sum()), flush=True)
