import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.nn.functional as F
import gnn_tableextraction_b200 as gte
from gnn_tableextraction_b200 import synth
from helpers import oracle_graph_from_pages
from oracle import sage_oracle as so
def rel(a,b):
    a,b=a.double().cpu(),b.double().cpu(); return ((a-b).abs().max()/b.abs().max()).item()
for kw in (dict(num_pages=5,k=5,bidirectional=True), dict(num_pages=8), dict(num_pages=5), dict(num_pages=5,k=5), dict(num_pages=12,k=5,bidirectional=True)):
    pages=synth.make_pages(**kw); og=oracle_graph_from_pages(pages)
    torch.manual_seed(0); om=so.OracleGcnSAGE(13,218,9,3,F.relu,0); cm=gte.GcnSAGE(13,218,9,3,F.relu,0); cm.load_state_dict(om.state_dict()); cm=cm.cuda()
    g=gte.PageGraphBatch.from_pages(pages,"cuda")
    logits=cm(g); ref=om(og)
    gte.cross_entropy(logits,g.ndata["label"]).backward()
    torch.nn.CrossEntropyLoss()(ref,og.ndata["label"].long()).backward()
    # fp64 oracle grads
    om64=so.OracleGcnSAGE(13,218,9,3,F.relu,0).double(); om64.load_state_dict({k:v.double() for k,v in om.state_dict().items()})
    og64=oracle_graph_from_pages(pages); og64.ndata["feat"]=og64.ndata["feat"].double(); og64.edata["feat"]=og64.edata["feat"].double()
    r64=om64(og64); torch.nn.CrossEntropyLoss()(r64,og.ndata["label"].long()).backward()
    print(kw, "mode", os.environ.get("GTE_GEMM","auto"), "N", g.num_nodes(), "logits cuda-vs-fp32oracle %.2e cuda-vs-fp64 %.2e oracle32-vs-fp64 %.2e"%(rel(logits,ref), rel(logits,r64), rel(ref,r64)))
    for (k,p),(_,q),(_,q64) in zip(cm.named_parameters(),om.named_parameters(),om64.named_parameters()):
        print("   %-28s cuda-vs-oracle32 %.2e | cuda-vs-fp64 %.2e | oracle32-vs-fp64 %.2e"%(k, rel(p.grad,q.grad), rel(p.grad,q64.grad), rel(q.grad,q64.grad)))
