import os

import torch

from tf_yarn_b200.pytorch import model_ckpt


def _model_and_opt():
    model = torch.nn.Linear(4, 2)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    model(torch.randn(3, 4)).sum().backward()
    opt.step()
    return model, opt


def test_save_and_load_round_trip(tmp_path):
    model, opt = _model_and_opt()
    path = model_ckpt.save_ckpt(str(tmp_path), model, opt, epoch=3, loss=0.25)
    assert os.path.basename(path) == "model_3.pt"
    model2 = torch.nn.Linear(4, 2)
    opt2 = torch.optim.SGD(model2.parameters(), lr=0.1, momentum=0.9)
    ckpt = model_ckpt.load_ckpt(path, model2, opt2, "cpu")
    assert ckpt["epoch"] == 3 and ckpt["loss"] == 0.25
    for a, b in zip(model.parameters(), model2.parameters()):
        assert torch.equal(a, b)
    assert opt2.state_dict()["state"].keys() == opt.state_dict()["state"].keys()


def test_find_and_load_latest(tmp_path):
    model, opt = _model_and_opt()
    assert model_ckpt.find_latest_ckpt(str(tmp_path)) is None
    assert model_ckpt.load_latest_ckpt(str(tmp_path), model, opt, "cpu") is None
    for epoch in (1, 12, 5):
        model_ckpt.save_ckpt(str(tmp_path), model, opt, epoch)
    (tmp_path / "notes.txt").write_text("not a checkpoint")
    assert model_ckpt.find_latest_ckpt(str(tmp_path)).endswith("model_12.pt")
    assert model_ckpt.load_latest_ckpt(str(tmp_path), model, opt, "cpu")["epoch"] == 12


def test_unwrap_model_strips_wrappers():
    inner = torch.nn.Linear(2, 2)

    class Wrapper(torch.nn.Module):
        _is_tfy_ddp = True

        def __init__(self, m):
            super().__init__()
            self.module = m
    assert model_ckpt._unwrap_model(Wrapper(inner)) is inner
    assert model_ckpt._unwrap_model(inner) is inner
