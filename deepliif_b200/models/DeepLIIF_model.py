"""DeepLIIFModel: modality generators G1..Gn, seg generators G{S}0..G{S}n and (training) their PatchGAN
discriminators (reference deepliif/models/DeepLIIF_model.py:8-507).

  forward (:175-203)   fake_B_i = G_i(real_A);  fake_B_S_0 = GS_0(real_A), fake_B_S_i = GS_i(fake_B_i);
                       fake_B_S = sum_i w_i * fake_B_S_i   (on-device dlb_seg_finish, list order, fp32)
Every network call runs on the sm_100a engines (models/networks.py)."""
import torch

from . import networks
from .base_model import BaseModel
from .. import ops
from ..util.util import init_input_and_mod_id


class DeepLIIFModel(BaseModel):
    def __init__(self, opt):
        BaseModel.__init__(self, opt)
        if not hasattr(opt, "net_gs"):
            opt.net_gs = "unet_512"
        self.seg_gen, self.seg_weights = opt.seg_gen, opt.seg_weights
        self.loss_G_weights, self.loss_D_weights = opt.loss_G_weights, opt.loss_D_weights
        self.mod_id_seg, self.input_id = init_input_and_mod_id(opt)
        print(f"Initializing model with segmentation modality id {self.mod_id_seg}, input id {self.input_id}")
        self.gpu_ids = opt.gpu_ids if opt.is_train else []      # test time: plain modules (no wrapper)
        n, S = opt.modalities_no, self.mod_id_seg
        base = 0 if self.input_id == "0" else 1
        self.visual_names = ["real_A"]
        for i in range(1, n + 1):
            self.loss_names += [f"G_GAN_{i}", f"G_L1_{i}", f"D_real_{i}", f"D_fake_{i}"]
            self.visual_names += [f"fake_B_{i}", f"real_B_{i}"]
        if self.seg_gen:
            self.loss_names += [f"G_GAN_{S}", f"G_L1_{S}", f"D_real_{S}", f"D_fake_{S}"]
            self.visual_names += [f"fake_B_{S}{i}" for i in range(n + 1)] + [f"fake_B_{S}", f"real_B_{S}"]
        self.model_names_g = [f"G{i}" for i in range(1, n + 1)]
        self.model_names_gs = [f"G{S}{i + base}" for i in range(n + 1)] if self.seg_gen else []
        self.model_names_d = [f"D{i}" for i in range(1, n + 1)] if self.is_train else []
        self.model_names_ds = [f"D{S}{i + base}" for i in range(n + 1)] if (self.is_train and self.seg_gen) else []
        self.model_names = self.model_names_g + self.model_names_gs + self.model_names_d + self.model_names_ds
        if isinstance(opt.netG, str):
            opt.netG = [opt.netG] * n
        if isinstance(opt.net_gs, str):
            opt.net_gs = [opt.net_gs] * (n + 1)
        in_nc = opt.input_nc * opt.input_no
        for i, name in enumerate(self.model_names_g):
            setattr(self, "net" + name, networks.define_G(in_nc, opt.output_nc, opt.ngf, opt.netG[i], opt.norm,
                                                          not opt.no_dropout, opt.init_type, opt.init_gain, self.gpu_ids,
                                                          opt.padding, getattr(opt, "upsample", "convtranspose")))
        for i, name in enumerate(self.model_names_gs):       # seg generators use define_G's default padding (reflect)
            setattr(self, "net" + name, networks.define_G(in_nc, opt.output_nc, opt.ngf, opt.net_gs[i], opt.norm,
                                                          not opt.no_dropout, opt.init_type, opt.init_gain, self.gpu_ids))
        if self.is_train:
            from .. import training
            training.install()
        for name in self.model_names_d + self.model_names_ds:
            setattr(self, "net" + name, networks.define_D(in_nc + opt.output_nc, opt.ndf, opt.netD, opt.n_layers_D,
                                                          opt.norm, opt.init_type, opt.init_gain, self.gpu_ids))
        if not self.is_train and torch.cuda.is_available():
            dev = torch.device("cuda", opt.gpu_ids[0] if opt.gpu_ids and opt.gpu_ids[0] >= 0 else torch.cuda.current_device())
            self.device = dev
            for name in self.model_names:
                self._net(name).to(dev)
        if self.is_train:
            self.criterionGAN_mod = networks.GANLoss(opt.gan_mode).to(self.device)
            self.criterionGAN_seg = networks.GANLoss(opt.gan_mode_s).to(self.device)
            self.criterionSmoothL1 = torch.nn.SmoothL1Loss()
            g_params = [p for nm in self.model_names_g + self.model_names_gs for p in self._net(nm).parameters()]
            d_params = [p for nm in self.model_names_d + self.model_names_ds for p in self._net(nm).parameters()]
            Opt = networks.get_optimizer(opt.optimizer)
            try:
                self.optimizer_G = Opt(g_params, lr=opt.lr_g, betas=(opt.beta1, 0.999))
                self.optimizer_D = Opt(d_params, lr=opt.lr_d, betas=(opt.beta1, 0.999))
            except TypeError:
                self.optimizer_G, self.optimizer_D = Opt(g_params, lr=opt.lr_g), Opt(d_params, lr=opt.lr_d)
            self.optimizers += [self.optimizer_G, self.optimizer_D]

    def set_input(self, input):
        A = input["A"]
        self.real_A = torch.cat([a.to(self.device) for a in A], dim=1) if isinstance(A, list) else A.to(self.device)
        self.real_B_array = input.get("B", [])
        for i in range(min(self.opt.modalities_no, len(self.real_B_array))):
            setattr(self, f"real_B_{i + 1}", self.real_B_array[i].to(self.device))
        if self.opt.seg_gen and len(self.real_B_array) > self.opt.modalities_no:
            setattr(self, f"real_B_{self.mod_id_seg}", self.real_B_array[self.opt.modalities_no].to(self.device))
        self.image_paths = input.get("A_paths", [])

    def forward(self):
        n, S = self.opt.modalities_no, self.mod_id_seg
        for i in range(n):
            setattr(self, f"fake_B_{i + 1}", self._net(f"G{i + 1}")(self.real_A))
        if self.seg_gen:
            parts = []
            for i, name in enumerate(self.model_names_gs):
                src = self.real_A if i == 0 else getattr(self, f"fake_B_{i}")
                parts.append(self._net(name)(src))
                setattr(self, f"fake_B_{S}_{i}", parts[-1])
            if self.is_train and torch.is_grad_enabled():
                # differentiable form of the same weighted sum (3-channel images: autograd glue)
                seg = torch.stack([torch.mul(p, w) for p, w in zip(parts, self.seg_weights)]).sum(dim=0)
            else:
                seg, _, _ = ops.seg_finish([p.contiguous() for p in parts], self.seg_weights, want_u8=False, want_mask=False)
            setattr(self, f"fake_B_{S}", seg)

    def optimize_parameters(self):
        from .. import training
        training.deepliif_step(self)
