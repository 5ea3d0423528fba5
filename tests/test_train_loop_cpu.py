"""gan_control_b200.train (the train_generator.py counterpart) on CPU with the kernel stand-ins: a tiny vanilla config
trains, writes reference-format checkpoints + args.json, resumes from the file (iteration taken from its name, gt.py:181-185)
and the result loads through the inference front end (inference.py:110-149 layout)."""
import json
import os

import torch

from gan_control_b200 import train as T
from gan_control_b200.inference import Inference

CONFIG = {
    'model_config': {'vanilla': True, 'img_channels': 3, 'split_fc': False, 'marge_fc': False, 'latent_size': 32, 'size': 8,
                     'n_mlp': 2, 'channel_multiplier': 2, 'conv_transpose': True, 'g_noise_mode': 'normal'},
    'training_config': {'iter': 3, 'start_iter': 0, 'batch': 4, 'mini_batch': 4, 'sub_groups_dict': {}, 'r1': 1, 'g_reg_every': 4,
                        'd_reg_every': 16, 'lr_g': 0.002, 'lr_d': 0.002, 'g_moving_average': 10000, 'path_regularize': 2,
                        'path_batch_shrink': 2, 'mixing': 0.5, 'save_nets_interval': 2},
    'ckpt_config': {'enabled': False, 'ckpt': 'no_ckpt'},
}


def test_train_checkpoint_resume_and_inference(cpu_kernels, tmp_path):
    save_dir = str(tmp_path / 'run')
    logs = []
    step = T.train(json.loads(json.dumps(CONFIG)), save_dir=save_dir, device='cpu', act_dtype=torch.float32, use_graphs=False,
                   log_every=1, log=logs.append)
    assert len(logs) == 3 and all(torch.isfinite(p).all() for p in step.g.parameters())
    files = sorted(os.listdir(os.path.join(save_dir, 'checkpoint')))
    assert files == ['000002.pt', '000003.pt'] and os.path.exists(os.path.join(save_dir, 'args.json'))
    ckpt = torch.load(os.path.join(save_dir, 'checkpoint', '000003.pt'))
    assert set(ckpt) >= {'g', 'd', 'g_ema', 'g_optim', 'd_optim'}
    assert float(ckpt['g_optim']['state'][0]['step']) == 4.0          # 3 plain steps + the path-length step of iteration 0
    # resume: two more iterations from the file
    cfg = json.loads(json.dumps(CONFIG))
    cfg['ckpt_config'] = {'enabled': True, 'ckpt': os.path.join(save_dir, 'checkpoint', '000003.pt')}
    logs2 = []
    step2 = T.train(cfg, save_dir=save_dir, iters=2, device='cpu', act_dtype=torch.float32, use_graphs=False, log_every=1, log=logs2.append)
    assert [l.split(':')[0] for l in logs2] == ['iter 3', 'iter 4']
    assert float(step2.g_optim.t[0]) == 7.0                             # 4 + iterations 3 and 4 (4 is a path-length iteration)
    assert os.path.exists(os.path.join(save_dir, 'checkpoint', '000005.pt'))
    # the run directory is what the inference front end reads (latest checkpoint's g_ema)
    g, groups, latent_size, config, ckpt_iter = Inference.retrieve_model(save_dir, device='cpu', act_dtype=torch.float32)
    assert ckpt_iter == '000005' and latent_size == 32 and not groups
    for (k, a), (_, b) in zip(g.named_parameters(), step2.g_ema.named_parameters()):
        assert torch.equal(a, b.detach()), k


def test_snapshot_restore_undoes_steps(cpu_kernels):
    """`GanTrainStep.capture` runs warm-up iterations before recording the CUDA graphs; `_snapshot` / `_restore` must bring
    parameters, Adam state, EMA and the path-length mean back exactly (in place: modules keep viewing the arenas)."""
    from gan_control_b200.train_step import GanTrainStep
    torch.manual_seed(0)
    step = GanTrainStep.from_config(json.loads(json.dumps(CONFIG)), device='cpu', act_dtype=torch.float32, mixing=0)
    real = torch.randn(4, 3, 8, 8).clamp_(-1, 1)
    step.train_step(1, real)                                   # some non-trivial optimiser state first
    before = [t.clone() for t in step._state_tensors()]
    ptrs = [p.data_ptr() for p in step.g.parameters()]
    snap = step._snapshot()
    step.train_step(0, real)                                   # both regularisers: every state tensor changes
    assert any(not torch.equal(a, b) for a, b in zip(before, step._state_tensors()))
    step._restore(snap)
    assert all(torch.equal(a, b) for a, b in zip(before, step._state_tensors()))
    assert ptrs == [p.data_ptr() for p in step.g.parameters()]
    step.g_arena.check_views()


def test_graphed_step_returns_the_replayed_variants_losses(cpu_kernels):
    """train_step_graphed hands back the loss tensors of the graphs it replayed (each variant keeps its own): checked here
    with stand-in graph objects, the capture itself needs CUDA (scripts/train_smoke.py)."""
    from gan_control_b200.train_step import GanTrainStep
    step = GanTrainStep.from_config(json.loads(json.dumps(CONFIG)), device='cpu', act_dtype=torch.float32, mixing=0)
    replayed = []

    class FakeGraph:
        def __init__(self, name):
            self.name = name

        def replay(self):
            replayed.append(self.name)
    step.static_real = torch.zeros(4, 3, 8, 8)
    step.graphs = {n: FakeGraph(n) for n in ('d', 'd_reg', 'g', 'g_reg')}
    step.graph_launches = {n: 1 for n in step.graphs}
    step.replayed_launches = 0
    step.graph_out = {'d': {'d_loss': torch.tensor(1.0)}, 'd_reg': {'r1_loss': torch.tensor(2.0)},
                      'g': {'g_loss': torch.tensor(3.0)}, 'g_reg': {'g_loss': torch.tensor(4.0), 'path_loss': torch.tensor(5.0)}}
    real = torch.ones(4, 3, 8, 8)
    d0, g0 = step.train_step_graphed(0, real)
    d1, g1 = step.train_step_graphed(1, real)
    assert replayed == ['d', 'd_reg', 'g_reg', 'd', 'g'] and step.replayed_launches == 5
    assert (float(d0), float(g0), float(d1), float(g1)) == (1.0, 4.0, 1.0, 3.0)
    assert torch.equal(step.static_real, real)
