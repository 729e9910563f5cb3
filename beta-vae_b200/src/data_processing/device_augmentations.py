"""Device-side counterpart of the reference's ``data_processing.augmentations`` (src/data_processing/augmentations.py:5-28).

The reference's ``get_train_transforms()`` returns a torchvision ``Compose`` that runs per PIL image inside the
DataLoader workers: Resize -> [RandomHorizontalFlip] -> [RandomRotation(d)] -> [ColorJitter(brightness=b)] -> ToTensor.
That module is left alone (this package extends, not shadows, ``data_processing``), so the reference's own loader keeps
working.  The functions here return the batch-level stage the B200 loop uses instead: the loader ships sized uint8
batches, and flip / rotation / brightness / ToTensor's ``/ 255`` run in one kernel on the copy stream, with the same
result as torchvision for the same drawn parameters (tests/test_gpu_augment.py).  ``Resize`` of off-size source files
and image decoding stay on the host (dataset code, out of scope)."""
import os
import sys

_SRC = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (_SRC, os.path.dirname(_SRC)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from bvae_rt.augment import DeviceAugment
from utils.brain_tumor_utils.config_parser import get_config


def get_device_train_augment(generator=None, reference_order=True):
    """``DeviceAugment`` configured from ``cfg.augmentation`` exactly as get_train_transforms reads it
    (use_augmentations, horizontal_flip, rotation_degrees > 0, brightness > 0); None when augmentation is off."""
    return DeviceAugment.from_config(get_config(), generator=generator, reference_order=reference_order)


def get_device_test_augment():
    """The test transform is Resize + ToTensor only (augmentations.py:22-28): nothing to do on the device."""
    return None
