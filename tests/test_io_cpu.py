"""On-disk compatibility helpers, no GPU: the reference's PNG naming / channel order / quantisation (utils.py:116-167)."""
import os

import numpy as np
import torch


def test_save_imgs_naming_quantisation_and_read_back(tmp_path):
    from spaa_b200 import utils as ut
    x = torch.rand(3, 3, 6, 9, generator=torch.Generator().manual_seed(0))
    ut.save_imgs(x, str(tmp_path / "a"), idx=10)
    assert sorted(os.listdir(tmp_path / "a")) == ["img_0011.png", "img_0012.png", "img_0013.png"]       # numbered from idx + 1
    back = ut.torch_imread(str(tmp_path / "a" / "img_0012.png"))
    assert torch.equal(back, torch.from_numpy(np.uint8(x[1].numpy() * 255)).float() / 255)                  # truncation (np.uint8), RGB order kept
    u8 = (x.permute(0, 2, 3, 1) * 255).to(torch.uint8).numpy()                                              # uint8 [N,H,W,C] is written as is
    ut.save_imgs(u8, str(tmp_path / "b"))
    assert torch.equal(ut.torch_imread(str(tmp_path / "b" / "img_0001.png")), torch.from_numpy(u8[0]).permute(2, 0, 1).float() / 255)
    stack = ut.torch_imread_mt(str(tmp_path / "a"), index=[2, 0])
    assert stack.shape == (2, 3, 6, 9) and torch.equal(stack[1], ut.torch_imread(str(tmp_path / "a" / "img_0011.png")))
    gray = ut.torch_imread_mt(str(tmp_path / "a"), size=(3, 4), gray_scale=True, normalize=True)
    assert gray.shape == (3, 1, 3, 4) and gray.min() >= -1 and gray.max() <= 1


def test_attacker_folder_names_and_small_helpers():
    from spaa_b200 import utils as ut
    from spaa_b200.projector_based_attack import shard_jobs, to_attacker_cfg_str
    # the folder names the reference derives (projector_based_attack.py:194-210 with get_model_train_cfg defaults)
    assert to_attacker_cfg_str("SPAA") == ("SPAA_PCNet_l1+ssim_500_24_2000", "PCNet_l1+ssim_500_24_2000")
    assert to_attacker_cfg_str("PerC-AL+CompenNet++") == ("PerC-AL+CompenNet++_l1+ssim_500_24_2000", "CompenNet++_l1+ssim_500_24_2000")
    assert float(ut.l2_norm_to_mse(torch.full((2, 4, 4), 3.0), 3)) == 3.0
    assert ut.idx_to_label({0: "a", 5: "b", 9: "c"}, [2, 0]) == ["c", "a"]
    jobs = list("abcdefg")
    parts = [shard_jobs(jobs, r, 3) for r in range(3)]
    assert sorted(i for p in parts for i, _ in p) == list(range(7)) and parts[1] == [(1, "b"), (4, "e")]
