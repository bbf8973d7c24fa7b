import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, random, ctypes as C, os
from bayesformer_b200 import _lib, engine
from bayesformer_b200.libs import preprocessing
from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
from bayesformer_b200.model.transformer import MyTransformer
dev=torch.device("cuda:0"); tok=ReversedPairingTokenizer(); random.seed(1)
tr = preprocessing.create_dataset(100,3)
torch.manual_seed(0)
m = MyTransformer(tok.ntokens,64,8,8,2,bayes_dropout=0.3,bayes=True).to(dev)
trn = engine.FusedTrainer(m, 1e-3)
pr, an, p_len, a_len, _, _ = preprocessing.get_batch(tr, tok, 0, 20)
tokens = torch.cat((pr, an), 0).to(dev)
lib=_lib.load()
for _ in range(3): trn.step(tokens, p_len, 0.1, 0.3)
buf=(C.c_ulonglong*128)()
lib.bfb200_debug_train_timing.restype=C.c_int
lib.bfb200_debug_train_timing(buf)
trn.step(tokens, p_len, 0.1, 0.3)
n=lib.bfb200_debug_train_timing(buf)
ts=[buf[i] for i in range(n)]
print(n, "barriers; per-phase us:", [round((ts[i+1]-ts[i])/1e3,1) for i in range(n-1)])
print("total", (ts[-1]-ts[0])/1e3)
