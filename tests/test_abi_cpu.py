"""CPU-side checks (no GPU needed): the C-ABI library loads and exports every symbol the header
declares, the host-side mirror has the reference's parameter names/shapes, and the product path
refuses to run without CUDA instead of silently falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import dcnr_b200
from dcnr_b200 import _cabi as C
from tests.helpers import load_model_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "dcnr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcnr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(C.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    raw = ctypes.CDLL(C.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/dcnr.h but not exported"
    assert set(C.EXPORTS) == set(names), set(C.EXPORTS) ^ set(names)
    assert C.lib().dcnr_abi_version() == C.ABI_VERSION == 4


def test_ctypes_structs_match_header_sizes(tmp_path):
    """sizeof() of every struct as gcc sees include/dcnr.h must equal the ctypes mirror."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "dcnr.h"\nint main(void){printf("%zu %zu %zu %zu\\n", '
                   'sizeof(dcnr_dims), sizeof(dcnr_params), sizeof(dcnr_grads), sizeof(dcnr_batch));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(C.Dims), ctypes.sizeof(C.Params), ctypes.sizeof(C.Grads), ctypes.sizeof(C.Batch)]


def test_argument_errors_are_status_codes_not_crashes():
    L = C.lib()
    assert L.dcnr_knn_normalize(None, None, 0, 4, None) == C.ERR_INVALID
    assert b"null" in L.dcnr_last_error_string()
    d = C.Dims()
    assert L.dcnr_workspace_bytes(ctypes.byref(d), 16, 0) == -1      # hidden == 0 is rejected
    with pytest.raises(RuntimeError):
        C.check(C.ERR_INVALID)


@pytest.mark.parametrize("name", ["p0", "odd", "min"])
def test_state_dict_keys_and_shapes_match_reference(name):
    case = load_model_case(name)
    m = dcnr_b200.DCN_RecSys(case["n_users"], case["n_items"], case["cat_dims"], case["n_num"], case["params"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(case["state"].keys()) or set(sd.keys()) == set(case["state"].keys())
    for k, v in case["state"].items():
        assert tuple(sd[k].shape) == tuple(v.shape) and sd[k].dtype == v.dtype, k
    m.load_state_dict(case["state"])
    assert m.item_embedding.weight.shape == (case["n_items"], case["params"]["emb_dim"])   # read at train.py:393
    # stock optimizers accept .parameters() (train.py:201-204)
    torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-2)


def test_extra_param_keys_ignored_and_default_res_blocks():
    m = dcnr_b200.DCN_RecSys(5, 4, {"city": 3}, 2, dict(emb_dim=16, hidden_dim=32, n_cross_layers=2, dropout=0.1,
                                                        lr=1e-3, batch_size=512, optimizer_name="Adam"))
    assert len(m.res_blocks) == 2 and len(m.cross_network) == 2      # n_res_blocks default (train.py:134)
    assert m.cat_embeddings[0].weight.shape == (3, int(np.sqrt(3)) + 1)
    with pytest.raises(KeyError):
        dcnr_b200.DCN_RecSys(5, 4, {"city": 3}, 2, dict(emb_dim=16, hidden_dim=32, n_cross_layers=2))  # dropout required


def test_no_cpu_fallback():
    m = dcnr_b200.DCN_RecSys(5, 4, {"city": 3}, 2, dict(emb_dim=16, hidden_dim=32, n_cross_layers=1, dropout=0.0))
    z = torch.zeros(2, dtype=torch.long)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(z, z, torch.zeros(2, 1, dtype=torch.long), torch.zeros(2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        dcnr_b200.CrossLayer(8)(torch.zeros(2, 8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            dcnr_b200.NearestNeighbors().fit(np.zeros((4, 4), dtype=np.float32))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.dirname(C.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
