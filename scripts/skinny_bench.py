"""Narrow-operand dense ops of the input layer (K = 2*13) and the class layer (Fout = 9) at config-2 size:
tensor-core route vs CUDA-core route, CUDA events, 3 warm-up + 20 timed launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tableextraction_b200 import ops
DEV = "cuda"
n = int(os.environ.get("N", "153600"))
g = torch.Generator(device=DEV).manual_seed(0)
def rnd(r, c):
    t = ops.empty_padded(r, c, DEV); t.normal_(generator=g); return t
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
h13, ah13, y218, dz218 = rnd(n, 13), rnd(n, 13), rnd(n, 218), rnd(n, 218)
dz9a, dz9b = rnd(n, 9), rnd(n, 9)
W1 = torch.randn(218, 26, device=DEV, generator=g) * 0.1
b1 = torch.randn(218, device=DEV, generator=g); gam = torch.ones(218, device=DEV); bet = torch.zeros(218, device=DEV)
W3 = torch.randn(9, 436, device=DEV, generator=g) * 0.1
b3 = torch.randn(9, device=DEV, generator=g)
dW1, db1 = torch.empty_like(W1), torch.empty(218, device=DEV)
dW3, db3 = torch.empty_like(W3), torch.empty(9, device=DEV)
res = {}
# input layer forward: z, y = LN(relu) of [h|ah] W1^T + b
p1 = ops.umma_pack_weights(W1, 13, 2)
res["in_fwd umma(fused LN)"] = t(lambda: ops.umma_linear_fwd(h13, ah13, 13, p1, b1, 218, gamma=gam, beta=bet, eps=1e-5, relu=True, fuse_ln=True))
def ffma_in():
    z = ops.linear_fwd(h13, ah13, W1, b1); return ops.layernorm_act_fwd(z, gam, bet, 1e-5, True)
res["in_fwd ffma+LN"] = t(ffma_in)
# input layer dW (+db comes from LN bwd in the step: pass None)
res["in_dW umma"] = t(lambda: ops.umma_linear_bwd_weight(dz218, h13, ah13, dW1, None))
res["in_dW ffma(gram)"] = t(lambda: ops.linear_bwd_weight(dz218, h13, ah13, dW1, None))
# class layer forward (stacked)
p3 = ops.umma_pack_weights(W3, 218, 2)
res["cls_fwd umma stacked"] = t(lambda: ops.umma_linear_fwd_stacked(y218, 218, p3, b3, 9))
def ffma_cls():
    ops.linear_fwd(y218, None, W3, b3, w_col0=0); ops.linear_fwd(y218, None, W3, None, w_col0=218)
res["cls_fwd ffma x2"] = t(ffma_cls)
# class layer dW (two narrow gradients against one wide input) + db
res["cls_dW umma"] = t(lambda: ops.umma_linear_bwd_weight2(dz9a, dz9b, y218, dW3, 0, 218, db3))
res["cls_dW ffma(gram)"] = t(lambda: ops.linear_bwd_weight2(dz9a, dz9b, y218, dW3, 0, 218, db3))
# class layer input gradient
res["cls_dx umma"] = t(lambda: ops.umma_linear_bwd_data2(dz9a, dz9b, p3, 218))
res["cls_dx ffma"] = t(lambda: ops.linear_bwd_data2(dz9a, 0, dz9b, 218, W3, 218))
res["in_fwd wide_out(fused LN)"] = t(lambda: ops.wide_out(h13, ah13, W1.data_ptr(), W1.data_ptr() + 4 * 13, 1, 26, 218, b1, gamma=gam, beta=bet, eps=1e-5, relu=True, fuse_ln=True))
res["in_dW gram_stream"] = t(lambda: ops.gram_stream(dz218, h13, ah13, dW1, 26, 1, dW1[:, 13:], 26, 1))
res["cls_dW gram_stream"] = t(lambda: ops.gram_stream(y218, dz9a, dz9b, dW3, 1, 436, dW3[:, 218:], 1, 436, qsum=db3))
res["cls_dx wide_out"] = t(lambda: ops.wide_out(dz9a, dz9b, W3.data_ptr(), W3.data_ptr() + 4 * 218, 436, 1, 218))
for k, v in res.items(): print(f"{k:28s} {v:8.1f} us")
