"""Shared test helpers: golden-fixture loading and synthetic inputs."""
import os
from collections import OrderedDict

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_model_case(name):
    z = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    n_users, n_items, n_num, B, E, H, L, R = [int(v) for v in z["meta"]]
    cat_names = ["city", "hotel_type", "extra", "extra2"]
    cat_dims = OrderedDict((cat_names[i], int(n)) for i, n in enumerate(z["cat_dims"]))
    params = dict(emb_dim=E, hidden_dim=H, n_cross_layers=L, n_res_blocks=R, dropout=0.0)
    state = OrderedDict((k[4:], torch.from_numpy(z[k])) for k in z.files if k.startswith("sd::"))
    case = dict(
        n_users=n_users, n_items=n_items, n_num=n_num, B=B, cat_dims=cat_dims, params=params, state=state,
        user_ids=torch.from_numpy(z["user_ids"]), item_ids=torch.from_numpy(z["item_ids"]),
        cat=torch.from_numpy(z["cat"]), num=torch.from_numpy(z["num"]), labels=torch.from_numpy(z["labels"]),
        grad_logits=torch.from_numpy(z["grad_logits"]), logits_eval=torch.from_numpy(z["logits_eval"]),
        logits_train=torch.from_numpy(z["logits_train"]), logits_train_f64=torch.from_numpy(z["logits_train_f64"]),
        loss=float(z["loss"]),
        grads=OrderedDict((k[6:], torch.from_numpy(z[k])) for k in z.files if k.startswith("grad::")),
        lossgrads=OrderedDict((k[10:], torch.from_numpy(z[k])) for k in z.files if k.startswith("lossgrad::")),
        after=OrderedDict((k[7:], torch.from_numpy(z[k])) for k in z.files if k.startswith("after::")),
    )
    return case


def synth_inputs(n_users, n_items, cat_dims, n_num, B, seed=1234, zipf=False, device="cpu"):
    """Synthetic batch in the hackathon_augmented_data.csv tensor schema."""
    g = torch.Generator().manual_seed(seed)
    if zipf:
        r = torch.rand(B, generator=g, dtype=torch.float64)
        user_ids = (n_users ** r - 1).long().clamp_(0, n_users - 1)   # log-uniform ~ Zipf(1) head-heavy ids
        r = torch.rand(B, generator=g, dtype=torch.float64)
        item_ids = (n_items ** r - 1).long().clamp_(0, n_items - 1)
    else:
        user_ids = torch.randint(0, n_users, (B,), generator=g)
        item_ids = torch.randint(0, n_items, (B,), generator=g)
    cat = torch.stack([torch.randint(0, n, (B,), generator=g) for n in cat_dims.values()], dim=1)
    num = torch.rand(B, n_num, generator=g)
    labels = (torch.rand(B, generator=g) < 0.3).float()
    return tuple(t.to(device) for t in (user_ids, item_ids, cat, num, labels))
