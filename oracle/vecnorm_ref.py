"""numpy restatement of stable-baselines 2.8.0 `RunningMeanStd` (common/running_mean_std.py) and the normalisation of
`VecNormalize` (common/vec_env/vec_normalize.py) — TEST INFRASTRUCTURE.  stable-baselines is not vendored under
/root/reference (requirements.txt:117), so this follows its published algorithm: parity UNPINNED."""
import numpy as np


class RunningMeanStd(object):
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, 'float64')
        self.var = np.ones(shape, 'float64')
        self.count = epsilon

    def update(self, arr):
        batch_mean = np.mean(arr, axis=0)
        batch_var = np.var(arr, axis=0)
        batch_count = arr.shape[0]
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / (self.count + batch_count)
        self.mean, self.var, self.count = new_mean, m_2 / (self.count + batch_count), tot_count


def normalize(x, rms, epsilon=1e-8, clip=10.0):
    return np.clip((x - rms.mean) / np.sqrt(rms.var + epsilon), -clip, clip)
