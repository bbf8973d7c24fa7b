"""torchrun --nproc-per-node G tools/check_distributed.py : the sharded acquisition over G GPUs (NCCL all-gather of
candidate keys) selects exactly the indices of the single-GPU run."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bayesformer_b200 import engine, sharding
from bayesformer_b200.libs import preprocessing
from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
from bayesformer_b200.model.transformer import MyTransformer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c2_transformer.npz"))
model = MyTransformer(114, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True)
sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
sd["pos_enc.pe"] = model.pos_enc.pe.clone()
model.load_state_dict(sd)
model = model.eval().to(dev)
tok = ReversedPairingTokenizer()
random.seed(77)
pool = preprocessing.create_dataset(20011, 3)
picked = sharding.select_samples_distributed(model, tok, pool, dev, nb_samples=200, compute="dropout", mode="margin", seed=5)
X, _, P, A, _, _ = preprocessing.get_batch(pool, tok, 0, len(pool))
pack = engine.scoring_pack(model, dev, P, A + 1)
whole = engine.mc_score_transformer(pack, X.to(dev), A + 1, 20, "dropout", "margin", seed=5)
want = engine.unpack_keys(engine.topk_keys(whole.pool_scores, 200))[1].tolist()
ok = picked == want
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} sharded == unsharded on every rank: {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
