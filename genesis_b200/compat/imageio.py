"""Stand-in for `imageio` (third_party/pytorch_fid/fid_score.py:46 imports `imread` at module level, and train.py:38
imports that module): PIL-backed `imread`, enough for the FID script's image loading."""
import numpy as np


def imread(path):
    from PIL import Image
    return np.asarray(Image.open(path))
