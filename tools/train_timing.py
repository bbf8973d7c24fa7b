"""utils.train on the GPU: eager loop vs CUDA-graph-captured steps (SURVEY C5's fine-tune phase: 700 items, 5 epochs)."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesformer_b200.libs import preprocessing, utils
from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
from bayesformer_b200.model.transformer import MyTransformer
dev = torch.device("cuda:0")
random.seed(1); torch.manual_seed(1)
tok = ReversedPairingTokenizer()
tr, va = preprocessing.create_dataset(700, 3), preprocessing.create_dataset(100, 3)
for graphs in ("0", "1"):
    os.environ["BFB200_GRAPH_TRAIN"] = graphs
    model = MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).to(dev)
    utils.train(model, tr, va, 1, 20, 1e-3, tok.ntokens, tok, dev)   # warm-up (cuBLAS handles, allocator)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    utils.train(model, tr, va, 5, 20, 1e-3, tok.ntokens, tok, dev)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    steps = 5 * len(range(0, len(tr) - 1, 20))
    print(f"BFB200_GRAPH_TRAIN={graphs}: 5 epochs x 35 batches = {steps} steps in {dt:.2f} s ({dt / steps * 1e3:.2f} ms per step incl. validation and graph capture)")
