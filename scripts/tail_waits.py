"""Per-role cycle counters of cnn_tail_bwd_kernel / cnn_tail_fwd_kernel (csrc/cnn_tail.cu) at the bench shape."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import manager_for
import news_recommendation_mind_b200 as mr
from news_recommendation_mind_b200 import data, _lib

lib = _lib.load()
lib.mr_debug_tapgemm_counters.argtypes = [ctypes.c_void_p]
N = 14080; L = 32; E = 300; H = 150
man = manager_for("cnn", "lstm", 5, 50, L, E, H, 10, precision="bf16")
emb = mr.BERT_Embedding(man, vocab_size=30522).cuda()
enc = mr.CNN_Encoder(man).cuda()
ids_t, mask_t = data.make_news_table(51282, L)
g = torch.Generator().manual_seed(0)
pick = torch.randint(0, ids_t.shape[0], (N,), generator=g)
ids = ids_t[pick].cuda(); mask = mask_t[pick].cuda()
for _ in range(2):
    enc.encode_ids(emb, ids, mask).sum().backward()
torch.cuda.synchronize()
NL = 12
buf = torch.zeros(NL, 148, 4, 5, dtype=torch.int64, device="cuda")
lib.mr_debug_tapgemm_counters(ctypes.c_void_p(buf.data_ptr()))
enc.encode_ids(emb, ids, mask).sum().backward()
torch.cuda.synchronize()
lib.mr_debug_tapgemm_counters(None)
b = buf.double().cpu()
for i in range(NL):
    if b[i].abs().sum() == 0:
        continue
    print("launch %d: kernel cycles avg %.0f" % (i, b[i, :, 0, 4].mean()))
    for r in range(4):
        tot = b[i, :, r, 4].mean()
        print("   role %d: %8.0f %8.0f %8.0f %8.0f  of %.0f cycles (%5.1f %5.1f %5.1f %5.1f %%)" % (
            r, b[i, :, r, 0].mean(), b[i, :, r, 1].mean(), b[i, :, r, 2].mean(), b[i, :, r, 3].mean(), tot,
            100 * b[i, :, r, 0].mean() / tot, 100 * b[i, :, r, 1].mean() / tot, 100 * b[i, :, r, 2].mean() / tot, 100 * b[i, :, r, 3].mean() / tot))
