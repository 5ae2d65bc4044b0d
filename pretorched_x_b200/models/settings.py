"""Registry of per-model preprocessing / checkpoint settings (reference: pretorched/models/settings.py:20-44),
restricted to the model families this engine implements."""
from .resnet3d import pretrained_settings as resnet3d_settings
from .nonlocalnet import pretrained_settings as nonlocal_settings
from .resnet2d import pretrained_settings as resnet2d_settings
from .resnext3d import pretrained_settings as resnext3d_settings      # settings.py:18,36 of the reference

all_settings = [resnet2d_settings, resnet3d_settings, resnext3d_settings, nonlocal_settings]

model_names = []
pretrained_settings = {}
for _settings in all_settings:
    for _name, _rows in _settings.items():
        pretrained_settings[_name] = _rows
        model_names.append(_name)
