import sys, numpy as np, torch
sys.path[:0]=['.', 'joint-kg-recommender_b200']
import kgrec_b200 as K
from oracle import kg_oracle as O
torch.manual_seed(1)
for d in (100,128,200):
    m = K.TransEModel(False, d, 5003, 11)
    W = {k.replace("_embeddings.weight",""): v.detach().cpu().numpy() for k,v in m.state_dict().items()}
    q = np.arange(37) ; r = np.arange(37)%11
    lt = lambda x: torch.as_tensor(x, dtype=torch.long, device='cuda')
    full = m.evaluateTail(lt(q), lt(r)).cpu().numpy()
    want = O.transe_eval(W["ent"], W["rel"], q, r, False, "tail")
    err = np.abs(full-want)/np.maximum(want,1e-6)
    print(d, "max rel err", err.max(), "bad rows", np.unique(np.where(err>1e-3)[0])[:10], "bad cols", np.unique(np.where(err>1e-3)[1])[:20])
