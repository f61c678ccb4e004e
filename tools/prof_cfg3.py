import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snuffy_b200 import snuffy_multiclass as mod
d, h, K, r, depth, C = 768, 8, 200, 0.5, 4, 2
i_cls = mod.FCLayer(d, C)
attn = mod.MultiHeadedAttention(h, d)
ff = mod.PositionwiseFeedForward(d, 4 * d, "relu", 0.0)
layer = mod.EncoderLayer(d, copy.deepcopy(attn), copy.deepcopy(ff), C, 0.0, K, r)
m = mod.MILNet(i_cls, mod.BClassifier(mod.Encoder(layer, depth), C, d)).cuda().eval()
for p in m.parameters():
    if p.dim() > 1:
        torch.nn.init.xavier_normal_(p)
for l in m.b_classifier.encoder.layers:
    l.return_attn = False
x = torch.randn(1, 6000, d, device="cuda")
with torch.no_grad():
    for _ in range(4):
        m(x)
torch.cuda.synchronize()
