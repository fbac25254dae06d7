"""Experiment configuration singleton with the reference's command line (reference config.py:20-126).

Same contract as the reference: the namespace is built from ``sys.argv`` AT IMPORT TIME with
``parse_known_args`` (``--training-dataset`` and ``--model`` are required), post-processed, printed, and
imported everywhere as ``from config import config``.  Crop types: 'oct' (2-D image already on the en-face
grid), 'relative_2d' (2-D *features* bilinearly resized to the en-face grid), 'relative_2d_max' (same with
adaptive max pooling), 'none'.
"""
import argparse
import socket

_FLAGS = [
    # (flag, kwargs)
    ('--debug', dict(action='store_true')),
    ('--training-dataset', dict(type=str, required=True)),
    ('--version', dict(type=str, default=None)),
    ('--data-ratio', dict(type=float, default=1.0)),
    ('--early-stopping', dict(type=int, default=None)),
    ('--exec-test', dict(action='store_true', help='execution test')),
    ('--epochs', dict(type=int, default=40)),
    ('--batch-size', dict(type=int, default=8)),
    ('--val-batch-size', dict(type=int, default=1)),
    ('--virtual-batch-size', dict(type=int, default=1)),
    ('--compression', dict(type=int, default=8)),
    ('--learning-rate', dict(type=float, default=1e-1)),
    ('--fusion-modality', dict(type=str, default=None)),
    ('--crop', dict(type=str, default='oct')),
    ('--model', dict(type=str, default=None, required=True)),
    ('--model-weights', dict(type=str, default=None)),
    ('--suffix', dict(type=str, default='')),
    ('--force-mem-cache-release', dict(default='ReleaseMemCache')),
    ('--number-of-outputs', dict(type=int, default=1)),
    ('--filly-annotations', dict(type=str, default=None)),
    ('--gpus', dict(type=int, nargs='+', default=1)),
    ('--threads', dict(type=int, default=8)),
    ('--split-indices', dict(nargs='+', type=int, default=[0, 1, 2, 3, 4])),
    ('--legacy-path', dict(action='store_true')),
    ('--use-complementary', dict(action='store_true', help='Force use of complementary data')),
    ('--split-name', dict(type=str, default=None)),
    ('--base-channels', dict(type=int, default=64)),
    ('--mask-variant', dict(type=str, default='faf', choices=['vs_proj', 'sq_proj_dil', 'oct', 'faf'],
                            help='mask variant, only for VRC vessel segmentation')),
    ('--multiplier', dict(type=int, default=20, help='Multiplier for the training dataset size.')),
    ('--rotation-augmentation', dict(action='store_true', help='Use rotation augmentation.')),
    ('--local-server-name', dict(type=str, default='server', choices=['server', 'msc_server'])),
]

parser = argparse.ArgumentParser()
for _flag, _kw in _FLAGS:
    parser.add_argument(_flag, **_kw)
config, _ = parser.parse_known_args()

config.DEBUG = config.debug
config.models_path = f'./__server_train/{config.version}/'
_name = config.model.lower()
config.use_complementary = ('fusion' in _name) or ('2d' in _name) or config.use_complementary
config.file_to_copy = 'run.sh'
config.layers = [1, 1, 2, 4]

if socket.gethostname() in ['hemingway']:                       # the authors' workstation override
    print('Running in local machine')
    config.models_path = f'./__train/{config.version}/'
    if config.model_weights is not None:
        config.model_weights = config.model_weights.replace('../', f'/mnt/Data/SSHFS/{config.local_server_name}/GA_SEG/')
    config.batch_size, config.gpus, config.split_indices = 1, [0], [0]
    config.virtual_batch_size, config.threads, config.multiplier = 1, 1, 1
    config.force_mem_cache_release = 'ReleaseMemCache'
    config.layers = [1, 1, 1, 1]

config.number_of_channels = [int(32 * 2 ** i) for i in range(len(config.layers))]

print('-' * 80)
print('[config]')
for _k, _v in config.__dict__.items():
    print(f'{_k}: {_v}')
print('-' * 80)
