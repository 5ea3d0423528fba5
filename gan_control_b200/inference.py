"""Inference front end on the B200 path: the reference's `Inference` / `Controller` call surface
(`inference/inference.py:17-149`, `inference/controller.py:15-116`, notebook cells 3-35) over
`gan_control_b200.modules.Generator` -- SURVEY.md §8(f) row 1.

Same method names, arguments and return values (`gen_batch`, `gen_batch_by_controls`,
`insert_group_w_latent`, `get_group_w_latent`, `calc_mean_w_latents`, `reset_noise`,
`expend_noise`), same on-disk layout (`<dir>/args.json`, `<dir>/checkpoint/<iter>.pt` with a `g_ema`
entry; controllers in `<dir>/<group>*/` with a `controller` entry).  Differences: no `nn.DataParallel`
wrapper (one process per GPU), the eval forward runs without autograd so the mapping network is the single
persistent kernel, activations are bf16 channels-last by default.

Fast path: without autograd every layer runs as `ops.mod_conv` (weight (de)modulation in one
kernel + one convolution kernel with the fused noise / bias / activation epilogue, nothing saved for a backward pass),
ToRGB carries its bias in the convolution epilogue, and the whole synthesis network is captured ONCE per batch size in a
CUDA graph (`cuda_graphs=True`, CUDA only) that later calls replay: ~70 launches become one graph launch.
"""
import json
import os

import torch

from . import modules as M


def _latest_checkpoint(model_dir):
    ckpts = sorted(os.listdir(os.path.join(model_dir, 'checkpoint')))
    return os.path.join(model_dir, 'checkpoint', ckpts[-1]), ckpts[-1].split('.')[0]


class Inference:
    def __init__(self, model_dir=None, *, generator=None, sub_groups_dict=None, latent_size=512, device='cuda',
                 act_dtype=torch.bfloat16, cuda_graphs=True):
        self.device = torch.device(device)
        self.cuda_graphs = bool(cuda_graphs) and self.device.type == 'cuda'
        self._graphs = {}
        if model_dir is not None:
            generator, sub_groups_dict, latent_size, self.config, self.ckpt_iter = self.retrieve_model(
                model_dir, device=self.device, act_dtype=act_dtype)
        self.model_dir = model_dir
        self.model = generator.to(self.device).eval()
        self.latent_size = latent_size
        self.sub_groups_dict = sub_groups_dict or {}
        # MiniBatchUtils.get_ordered_group_names / place_in_latent_dict (mini_batch_multi_split_utils.py:32-55)
        self.place_in_latent_dict = {k: list(v['place_in_latent']) for k, v in self.sub_groups_dict.items()}
        self.sub_group_names = sorted(self.place_in_latent_dict, key=lambda k: self.place_in_latent_dict[k][0])
        self.noise = None
        self.reset_noise()
        self.mean_w_latent = None
        self.mean_w_latents = None

    # -- model loading (inference.py:110-149) -----------------------------------------------------
    @staticmethod
    def retrieve_model(model_dir, device='cuda', act_dtype=torch.bfloat16):
        with open(os.path.join(model_dir, 'args.json')) as f:
            config = json.load(f)
        mc, tc = config['model_config'], config['training_config']
        ckpt_path, ckpt_iter = _latest_checkpoint(model_dir)
        ckpt = torch.load(ckpt_path, map_location='cpu')
        groups = None if mc.get('vanilla') else tc['sub_groups_dict']
        fc_config = M.FcConfig.from_sub_groups_dict(groups) if groups else None
        g = M.Generator(mc['size'], mc['latent_size'], mc['n_mlp'], channel_multiplier=mc['channel_multiplier'],
                        out_channels=mc['img_channels'], split_fc=mc['split_fc'], fc_config=fc_config,
                        conv_transpose=mc['conv_transpose'], noise_mode=mc.get('g_noise_mode', 'normal'),
                        act_dtype=act_dtype)
        g.load_state_dict(ckpt['g_ema'])
        return g.to(device), groups, mc['latent_size'], config, ckpt_iter

    # -- mean latents for truncation (inference.py:27-40) -----------------------------------------------
    @torch.no_grad()
    def calc_mean_w_latents(self, n_batches=100, batch=1000):
        acc = torch.zeros(self.latent_size, device=self.device)
        for _ in range(n_batches):
            z = torch.randn(batch, self.latent_size, device=self.device)
            acc += self.model.map_styles(z).mean(0)
        self.mean_w_latent = (acc / n_batches).cpu()
        self.mean_w_latents = {k: self.mean_w_latent[lo:hi] for k, (lo, hi) in self.place_in_latent_dict.items()}

    def reset_noise(self):
        self.noise = self.model.make_noise(device=self.device)

    # -- synthesis network, CUDA-graphed per batch size ---------------------------------------------------------------
    @torch.no_grad()
    def synthesize(self, latent, noise=None):
        """image = synthesis(W+ latent (B, n_latent, style_dim) or w (B, style_dim), noise list | None = fresh noise)
        (`Generator.forward` after the mapping network, gm.py:754-801)."""
        g = self.model
        if latent.ndim == 2:
            latent = latent.unsqueeze(1).repeat(1, g.n_latent, 1)
        if not self.cuda_graphs:
            return g.synthesis(latent, noise if noise is not None else [None] * g.num_layers)
        batch = latent.shape[0]
        entry = self._graphs.get(batch)
        if entry is None:
            st_latent = torch.zeros(batch, g.n_latent, g.style_dim, device=self.device)
            st_noise = [torch.zeros(batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), device=self.device)
                        for i in range(g.num_layers)]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                       # warm-up outside the capture (lazy allocations, tables)
                g.synthesis(st_latent, st_noise)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st_out = g.synthesis(st_latent, st_noise)
            entry = self._graphs[batch] = (graph, st_latent, st_noise, st_out)
        graph, st_latent, st_noise, st_out = entry
        st_latent.copy_(latent)
        for dst, src in zip(st_noise, noise if noise is not None else [None] * len(st_noise)):
            if src is None:
                dst.normal_()                                   # NoiseInjection's fresh draw (gm.py:343)
            else:
                dst.copy_(src.expand_as(dst))
        graph.replay()
        return st_out.clone()

    @staticmethod
    def expend_noise(noise, batch_size):
        return [n.repeat(batch_size, 1, 1, 1) for n in noise]

    def check_valid_group(self, group):
        if group not in self.sub_group_names:
            raise ValueError('group: %s not in valid group names for this model\nValid group names are:\n%s'
                             % (group, str(self.sub_group_names)))

    @torch.no_grad()
    def gen_batch(self, batch_size=1, normalize=True, latent=None, input_is_latent=False, static_noise=True,
                  truncation=1, **kwargs):
        if truncation < 1 and self.mean_w_latents is None:
            self.calc_mean_w_latents()
        injection_noise = None
        if latent is None:
            latent = torch.randn(batch_size, self.latent_size, device=self.device)
        elif input_is_latent:
            latent = latent.to(self.device)
            for group_key, value in kwargs.items():
                self.check_valid_group(group_key)
                if isinstance(value, str) and value == 'random':
                    # inference.py:66-68 indexes `[:, lo, lo]` (an IndexError on a 2-D w); the evident intent --
                    # resample that group's slice of w -- is what is done here
                    lo, hi = self.place_in_latent_dict[group_key]
                    w_rand = self.model.map_styles(torch.randn(latent.shape[0], self.latent_size, device=self.device))
                    latent = latent.clone()
                    latent[..., lo:hi] = w_rand[:, lo:hi] if latent.ndim == 2 else w_rand[:, None, lo:hi]
        latent = latent.to(self.device)
        if static_noise:
            self.reset_noise()
            injection_noise = self.expend_noise(self.noise, latent.shape[0])
        if truncation < 1:
            if not input_is_latent:
                latent = self.model.map_styles(latent)
                input_is_latent = True
            latent = latent.clone()
            for key, (lo, hi) in self.place_in_latent_dict.items():
                mean = self.mean_w_latents[key].to(self.device)
                latent[..., lo:hi] = truncation * (latent[..., lo:hi] - mean) + mean
        latent_w = latent if input_is_latent else self.model.map_styles(latent)
        if latent_w.ndim == 2:
            latent_w = latent_w.unsqueeze(1).repeat(1, self.model.n_latent, 1)       # gm.py:757-760
        tensor = self.synthesize(latent_w, injection_noise)
        if normalize:
            tensor = tensor.float().mul(0.5).add(0.5).clamp(min=0., max=1.).cpu()
        return tensor, latent, latent_w


class Controller(Inference):
    """`fc_controls`: {group or 'expression_q': FcStack}.  From a directory: `<dir>/generator` + one sub-directory
    per controlled group (controller.py:16-27,91-116)."""

    def __init__(self, controller_dir=None, *, fc_controls=None, **kw):
        if controller_dir is not None:
            super().__init__(os.path.join(controller_dir, 'generator'), **kw)
            self.fc_controls = {}
            for name in self.sub_group_names + ['expression_q']:
                ctl = self.retrieve_controller(controller_dir, name)
                if ctl is not None:
                    self.fc_controls[name] = ctl
        else:
            super().__init__(None, **kw)
            self.fc_controls = {k: v.to(self.device).eval() for k, v in (fc_controls or {}).items()}

    @staticmethod
    def get_controller_dir(controller_dir, sub_group_name):
        for d in sorted(os.listdir(controller_dir)):
            if d.startswith(sub_group_name) and not (sub_group_name == 'expression' and d.startswith('expression_q')):
                return os.path.join(controller_dir, d)
        return None

    def retrieve_controller(self, controller_dir, sub_group_name):
        path = self.get_controller_dir(controller_dir, sub_group_name)
        if path is None:
            return None
        with open(os.path.join(path, 'args.json')) as f:
            mc = json.load(f)['model_config']
        ckpt_path, _ = _latest_checkpoint(path)
        lo, hi = self.place_in_latent_dict['expression' if sub_group_name == 'expression_q' else sub_group_name]
        ctl = M.FcStack(mc['lr_mlp'], mc['n_mlp'], mc['in_dim'], mc['mid_dim'], hi - lo)
        ctl.load_state_dict(torch.load(ckpt_path, map_location='cpu')['controller'])
        return ctl.to(self.device).eval()

    def check_if_group_has_control(self, group):
        if group not in self.fc_controls:
            raise ValueError('group: %s has no control' % group)
        return True

    @torch.no_grad()
    def gen_batch_by_controls(self, batch_size=1, latent=None, normalize=True, input_is_latent=False, static_noise=True,
                              **kwargs):
        if latent is None:
            latent = torch.randn(batch_size, self.latent_size, device=self.device)
        latent = latent.clone().to(self.device)
        latent_w = latent if input_is_latent else self.model.map_styles(latent)
        for group_key, value in kwargs.items():
            if self.check_if_group_has_control(group_key):
                value = value.to(self.device).float()
                ctl = self.fc_controls['expression_q'] if (group_key == 'expression' and value.shape[1] == 8) \
                    else self.fc_controls[group_key]
                latent_w = self.insert_group_w_latent(latent_w, ctl(value), group_key)
        injection_noise = self.expend_noise(self.noise, latent.shape[0]) if static_noise else None
        tensor = self.synthesize(latent_w, injection_noise)
        if normalize:
            tensor = tensor.float().mul(0.5).add(0.5).clamp(min=0., max=1.).cpu()
        return tensor, latent, latent_w

    def generate_group_w_latent(self, group_key, value):
        return self.fc_controls[group_key](value)

    def insert_group_w_latent(self, latent_w, group_w_latent, group):
        lo, hi = self.place_in_latent_dict[group]
        if latent_w.ndim == 3:
            latent_w[:, :, lo:hi] = group_w_latent if group_w_latent.ndim == 3 else group_w_latent[:, None]
        else:
            latent_w[:, lo:hi] = group_w_latent
        return latent_w

    def get_group_w_latent(self, latent_w, group):
        lo, hi = self.place_in_latent_dict[group]
        return latent_w[..., lo:hi]
