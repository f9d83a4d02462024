"""CPU numerics model: which of the three f16x3 products can the FeedForward GEMMs (57 % of the FLOPs) drop without touching
code indices or the pixel bar?  Every other tensor-core product keeps the full form.  """
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import omni_oracle as oo
from util import load_golden, golden_setup, check_sub

def split(x):
    hi = x.clamp(-65504, 65504).half().float()
    lo = ((x - hi) * 2048).half().float()
    return hi, lo
def full(a, b):
    ah, al = split(a); bh, bl = split(b)
    return ah @ bh + (ah @ bl + al @ bh) / 2048
def make(ff1, ff2):
    def mm(a, b):
        mode = "full"
        if b.shape[-2:] == (512, 2730): mode = ff1
        elif b.shape[-2:] == (1365, 512): mode = ff2
        ah, al = split(a); bh, bl = split(b)
        if mode == "full": return ah @ bh + (ah @ bl + al @ bh) / 2048
        if mode == "hh": return ah @ bh
        if mode == "w_full": return ah @ bh + (ah @ bl) / 2048          # activations 11 bits, weights exact
        if mode == "a_full": return ah @ bh + (al @ bh) / 2048          # activations exact, weights 11 bits
    return mm
names = sys.argv[1:] or ["vid9x128_b2", "img256_cfg1"]
for name in names:
    fx = load_golden(name); cfg, sd, x = golden_setup(fx); is_image = x.ndim == 4
    for ff1, ff2 in (("full", "full"), ("hh", "hh"), ("a_full", "a_full"), ("w_full", "w_full"), ("a_full", "full"), ("full", "a_full"), ("hh", "full"), ("full", "hh")):
        with torch.no_grad():
            oo.MATMUL_MODEL = make(ff1, ff2)
            emb, idx = oo.encode(sd, cfg, x, include_embeddings=True)
            rec = oo.decode(sd, cfg, fx["idx"].long(), is_image)
            zerr = float((emb - fx["emb"]["full"]).abs().max()) if "full" in fx["emb"] else float("nan")
            print(f"{name:13s} FF1={ff1:7s} FF2={ff2:7s} flips {int((idx != fx['idx'].long()).sum())}/{idx.numel()}  max|dpx| {check_sub(fx['rec'], rec, 1.0, 'rec'):.2e}  max|dz_st| {zerr:.2e}", flush=True)
oo.MATMUL_MODEL = None
