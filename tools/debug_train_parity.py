import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.nn.functional as F
import cases as K, synth
import spgan_b200.functional as SF
from spgan_b200.discriminator import Discriminator
from spgan_b200.generator import Generator

def cerr(g, key, got):
    got = np.asarray(got)
    if key in g:
        return K.rel_err(got, g[key]), 0.0
    ref = g[key + "__sample"]; stride = -(-got.size // 4096); sample = got.reshape(-1)[::stride]
    norm = float(np.sqrt((got.astype(np.float64) ** 2).sum())); rn = float(g[key + "__norm"])
    scale = max(np.abs(ref).max(), rn / np.sqrt(got.size), 1e-30)
    l2 = float(np.sqrt(((sample - ref).astype(np.float64) ** 2).sum()) / max(np.sqrt((ref.astype(np.float64) ** 2).sum()), 1e-30))
    frac = float((np.abs(sample - ref) > 1e-3 * scale).mean())
    return "max %.2e  l2rel %.2e  frac>1e-3: %.4f  norm %.1e" % (float(np.abs(sample - ref).max() / scale), l2, frac, abs(norm - rn) / rn)

dev = torch.device("cuda:0")
for prec in (0, 1):
    SF.set_precision(prec)
    print("=== precision", prec)
    g = K.load("discriminator.npz")
    disc = Discriminator(); disc.load_state_dict(K.discriminator_state_dict()); disc = disc.to(dev).train()
    img = synth.randn_t(K.SEED, "d_img", (2, 3, 101, 101)).clamp(-1, 1).to(dev).requires_grad_(True)
    out = disc(img); d, ac = out["d_patch"], out["ac_coords_pred"]
    params = dict(disc.named_parameters())
    loss = F.softplus(-d).mean() + (ac * synth.randn_t(K.SEED, "d_acw", ac.shape).to(dev)).sum()
    grads = torch.autograd.grad(loss, [img] + [params[n] for n in K.D_GRAD_KEYS], retain_graph=True)
    print("D g_img", cerr(g, "g_img", K.t2n(grads[0])))
    for n, got in zip(K.D_GRAD_KEYS, grads[1:]):
        print("D", n, cerr(g, "g_" + n, K.t2n(got)))
    g = K.load("generator_train.npz")
    gen = Generator(); gen.load_state_dict(K.generator_state_dict()); gen = gen.to(dev).train()
    gl, lat, coords, cps, noises, go = K.generator_train_case()
    lat = lat.to(dev).requires_grad_(True)
    im = gen(gl.to(dev), lat, coords.to(dev), cps, noises=[n.to(dev) for n in noises], inject_index=5)
    print("G img", cerr(g, "img", K.t2n(im)))
    params = dict(gen.named_parameters())
    grads = torch.autograd.grad((im * go.to(dev)).sum(), [lat] + [params[k] for k in K.TRAIN_GRAD_KEYS])
    print("G g_lat", cerr(g, "g_lat", K.t2n(grads[0])))
    for k, got in zip(K.TRAIN_GRAD_KEYS, grads[1:]):
        print("G", k, cerr(g, "g_" + k, K.t2n(got)))
