import torch, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import nets
from deepliif_b200 import engine_train
from test_training_gpu import _rand, _leafify, _cmp
torch.set_num_threads(32)
for padding in ("zero", "reflect"):
    for hw in (64, 160):
        cfg = dict(n_blocks=2, norm="batch", use_dropout=False, padding_type=padding)
        sd = nets.make_state_dict(nets.resnet_param_shapes(3, 3, 64, 2, "batch", False, padding), 7, "stress")
        x = _rand((2, 3, hw, hw), 70).requires_grad_(True); dY = _rand((2, 3, hw, hw), 71)
        leaf = _leafify(sd)
        y_ref = nets.resnet_forward(x, leaf, norm_mode="batch", **cfg); (y_ref * dY).sum().backward()
        eng = engine_train.ResnetTrainEngine(sd, norm_mode="batch", precision="bf16x3", **cfg)
        y, ctx = eng.forward_train(x.detach().cuda())
        grads, dx = eng.backward(ctx, dY.cuda(), need_dx=True)
        worst = _cmp(grads, leaf, tol_l2=1.0)
        print(padding, hw, "fwd", (y.cpu() - y_ref.detach()).abs().max().item(), "worst", worst, "dx", ((dx.cpu() - x.grad).norm() / x.grad.norm()).item(), flush=True)
