import sys, torch, numpy as np
sys.path.insert(0, '.')
from maskdit_b200.vae import AutoencoderKLDecoder
from oracle import vae_oracle as VO
g = {k: torch.from_numpy(v) for k, v in np.load('tests/golden/vae_decode.npz').items()}
vae = AutoencoderKLDecoder(); vae.load_state_dict(VO.make_vae_state_dict(3)); vae = vae.cuda().eval()
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
z = g['z'].cuda()
a = vae.decode(z); b = vae.decode(z)
print('run-to-run', rel(a, b), 'vs golden', rel(a.cpu(), g['images']))
for mr in (1 << 21, 8192, 4096, 2048, 512, 64):
    vae.max_rows = mr
    c = vae.decode(z)
    print('max_rows', mr, 'rel vs unchunked', rel(c, a), 'vs golden', rel(c.cpu(), g['images']), 'finite', bool(torch.isfinite(c).all()))
try:
    vae(z, 'encode'); print('no raise!')
except NotImplementedError as e:
    print('raises ok')
