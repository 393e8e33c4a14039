"""The reference's realistic joint-limit classifier (envs/env.py:39 `load_model('realistic_arm_limits_model.h5')`,
agents/human.py:134-152): a 4 -> 64 -> 64 -> 64 -> 1 tanh MLP with a sigmoid output that says whether a shoulder (3 angles) /
elbow configuration is one a person can reach.  The weights are compiled out of the Keras file by tools/compile_assets.py
(this image has neither keras nor h5py); `predict_classes` is Keras's: output > 0.5."""
import os

import numpy as np

ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets', 'realistic_arm_limits_model.agmlp.npz')

_ACT = {'tanh': np.tanh, 'sigmoid': lambda z: 1.0 / (1.0 + np.exp(-z)), 'linear': lambda z: z, 'relu': lambda z: np.maximum(z, 0.0)}


class ArmLimitsModel:
    def __init__(self, path=ASSET):
        z = np.load(path)
        self.layers = [(z['W%d' % k].astype(np.float32), z['b%d' % k].astype(np.float32), str(z['act%d' % k])) for k in range(int(z['n_layers']))]

    def predict(self, x):
        """[n][4] float -> [n][1] probabilities (fp32 arithmetic, as the Keras model)"""
        a = np.atleast_2d(np.asarray(x, dtype=np.float32))
        for W, b, act in self.layers:
            a = _ACT[act](a @ W + b).astype(np.float32)
        return a

    def predict_classes(self, x):
        return (self.predict(x) > 0.5).astype(np.int32)


_model = None


def load_model():
    global _model
    if _model is None:
        _model = ArmLimitsModel()
    return _model
