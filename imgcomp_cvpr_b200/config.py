"""Configs of the hot path.

The reference reads its hyper-parameters with ``fjcommon.config_parser.parse``
(code/val.py:71-72, code/train.py:65-66) from files in a small language:

    use <relative path>          inherit another file            (ae_configs/cvpr/low:1)
    constrain name :: A, B       enum: ``name`` must be A or B   (ae_configs/base:3)
    key = <python expression>    assignment                      (ae_configs/base:1)

fjcommon is not in the reference tree; ``parse`` below restates that behaviour
so a val.py-style driver can keep pointing at the user's own config files.  The
effective values of the published configs (ae_configs/cvpr/{low,med,hi},
pc_configs/cvpr/{res_shallow,res_shallow_64}) are also available without any
file through ``ae_config(name)`` / ``pc_config(name)``.
"""
import os


class Config(object):
    """Attribute bag, like fjcommon's parsed config object."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def all_params_and_values(self):
        return sorted(self.__dict__.items())

    def __repr__(self):
        return 'Config(%s)' % ', '.join('%s=%r' % kv for kv in self.all_params_and_values())


class _Enum(str):
    pass


def _parse_into(path, values, constraints):
    with open(path) as f:
        lines = f.readlines()
    for raw in lines:
        line = raw.split('#', 1)[0].strip()
        if not line:
            continue
        if line.startswith('use '):
            _parse_into(os.path.normpath(os.path.join(os.path.dirname(path), line[4:].strip())),
                        values, constraints)
        elif line.startswith('constrain '):
            name, allowed = line[len('constrain '):].split('::')
            constraints[name.strip()] = [a.strip() for a in allowed.split(',')]
        else:
            key, expr = (s.strip() for s in line.split('=', 1))
            env = {a: _Enum(a) for allowed in constraints.values() for a in allowed}
            env.update(values)
            val = eval(expr, {'__builtins__': {}}, env)      # noqa: S307 -- config files are trusted user input
            if key in constraints:
                if str(val) not in constraints[key]:
                    raise ValueError('%s: %s = %r not in %s' % (path, key, val, constraints[key]))
                val = str(val)
            values[key] = val


def parse(path):
    """-> (Config, rel_path) like fjcommon.config_parser.parse; rel_path is the
    path below the nearest ``*_configs`` directory (used in log-dir names,
    code/logdir_helpers.py:130-151)."""
    values, constraints = {}, {}
    _parse_into(path, values, constraints)
    parts = os.path.abspath(path).split(os.sep)
    idx = max([i for i, p in enumerate(parts) if p.endswith('_configs')] or [len(parts) - 2])
    return Config(**values), os.sep.join(parts[idx + 1:])


# Effective values after ``use`` inheritance ("Config / flags").
_AE_BASE = dict(
    arch='CVPR', num_chan_bn=32, heatmap=True, normalization='FIXED', regularization_factor=0.005,
    arch_param_B=5, num_centers=6, centers_initial_range=(-2, 2), regularization_factor_centers=0.1,
    beta=500, H_target=1.2, distortion_to_minimize='ms_ssim', K_ms_ssim=5000, K_psnr=100,
    crop_size=(160, 160), batch_size=30, lr_initial=8e-5, lr_schedule='DECAY',
    lr_schedule_decay_interval=2, lr_schedule_decay_rate=0.1, lr_schedule_decay_staircase=True,
    optimizer='ADAM', optimizer_momentum=0.9, lr_centers_factor=None,
    train_autoencoder=True, train_probclass=True)
_AE = {
    'cvpr/low': dict(num_chan_bn=32, H_target=2 * 0.2),     # ae_configs/cvpr/low:3-4
    'cvpr/med': dict(num_chan_bn=32, H_target=2 * 0.6),     # ae_configs/cvpr/med:3-4
    'cvpr/hi': dict(num_chan_bn=64, H_target=1.0),          # ae_configs/cvpr/hi:3-4
}
_PC_BASE = dict(
    arch='res_shallow', kernel_size=3, arch_param__k=24, arch_param__non_linearity='relu',
    arch_param__fc=64, regularization_factor=None, learn_pad_var=False, use_centers_for_padding=True,
    lr_initial=1e-4, optimizer='ADAM', optimizer_momentum=0.9, lr_schedule='DECAY',
    lr_schedule_decay_interval=2, lr_schedule_decay_rate=0.1, lr_schedule_decay_staircase=True)
_PC = {
    'cvpr/res_shallow': dict(),
    'cvpr/res_shallow_64': dict(arch_param__k=64),          # pc_configs/cvpr/res_shallow_64:8
}


def ae_config(name='cvpr/low'):
    name = {'cvpr/high': 'cvpr/hi'}.get(name, name)          # README.md:133 says "high", the file is "hi"
    return Config(**dict(_AE_BASE, **_AE[name]))


def pc_config(name='cvpr/res_shallow'):
    return Config(**dict(_PC_BASE, **_PC[name]))
