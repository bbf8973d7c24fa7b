"""Drive the round-2 side kernels a few times each (for an ncu capture): the fused fine-tune step and the one-launch top-k."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesformer_b200 import engine
from bayesformer_b200.libs import preprocessing
from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
from bayesformer_b200.model.transformer import MyTransformer

dev = torch.device("cuda:0"); tok = ReversedPairingTokenizer(); random.seed(1); torch.manual_seed(0)
m = MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).to(dev)
trn = engine.FusedTrainer(m, 1e-3)
pr, an, p_len, a_len, _, _ = preprocessing.get_batch(preprocessing.create_dataset(100, 3), tok, 0, 20)
tokens = torch.cat((pr, an), 0).to(dev)
for _ in range(4):
    trn.step(tokens, p_len, 0.1, 0.3)
s = torch.rand(1_000_000, device=dev)
for _ in range(4):
    engine.topk_keys(s, 200)
torch.cuda.synchronize()
print("ok", float(trn.loss))
