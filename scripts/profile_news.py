"""Runs the bf16 news encoder (forward, optionally backward) at the bench shape a few times; used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import manager_for
import news_recommendation_mind_b200 as mr
from news_recommendation_mind_b200 import data

N = int(os.environ.get("PN", "14080")); L = 32; E = 300; H = 150
bwd = os.environ.get("PBWD", "0") == "1"
reps = int(os.environ.get("PREPS", "2"))
man = manager_for("cnn", "lstm", 5, 50, L, E, H, 10, precision="bf16")
emb = mr.BERT_Embedding(man, vocab_size=30522).cuda()
enc = mr.CNN_Encoder(man).cuda()
ids_t, mask_t = data.make_news_table(51282, L)
# the bench batch: 256 impressions x (5 candidates + 50 history slots, padded with news 0) = 14,080 titles
xb = data.make_train_batch(ids_t, mask_t, 256, 5, 50, seed=0)
ids = torch.cat([xb["cdd_encoded_index"].view(-1, L), xb["his_encoded_index"].view(-1, L)])[:N].cuda()
mask = torch.cat([xb["cdd_attn_mask"].view(-1, L), xb["his_attn_mask"].view(-1, L)])[:N].cuda()
for i in range(reps):
    if bwd:
        news = enc.encode_ids(emb, ids, mask)
        news.sum().backward()
    else:
        with torch.no_grad():
            news = enc.encode_ids(emb, ids, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5):
    if bwd:
        enc.encode_ids(emb, ids, mask).sum().backward()
    else:
        with torch.no_grad():
            enc.encode_ids(emb, ids, mask)
e1.record(); torch.cuda.synchronize()
print("ms per call: %.3f" % (e0.elapsed_time(e1) / 5))
