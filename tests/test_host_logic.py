"""CPU: host-side logic around the kernels that needs no device (schedules, flat-buffer layout, mesh grids)."""
import math

import torch


def test_cosine_lr_matches_torch_scheduler():
    """optim.cosine_lr == CosineAnnealingLR(T_max=100_000, eta_min=3e-4) of spurfies/train.py:191-193."""
    from spurfies_b200.optim import cosine_lr
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.Adam(p, lr=5e-4)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=100_000, eta_min=3e-4)
    for step in range(2000):
        if step % 250 == 0:
            assert math.isclose(cosine_lr(step), opt.param_groups[0]["lr"], rel_tol=1e-9), step
        opt.step()
        sch.step()


def test_flat_offsets_are_16_byte_aligned_and_ordered():
    from spurfies_b200.dist import FlatGradReducer, flat_offsets
    params = [torch.zeros(s) for s in [(5, 3), (7,), (), (4, 4)]]
    offs, total = flat_offsets(params, align=4)
    assert offs == [0, 16, 24, 28] and total == 44
    offs1, total1 = flat_offsets(params, align=1)
    assert offs1 == [0, 15, 22, 23] and total1 == 39
    red = FlatGradReducer([torch.nn.Parameter(p) for p in params], 1, align=4)
    flat = red.attach()
    assert red.attached() and flat.numel() == 44
    for p, o in zip(red.params, offs):
        assert p.grad.data_ptr() == flat.data_ptr() + 4 * o
