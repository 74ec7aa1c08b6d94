"""Grid Adaptive Curriculum with device-resident state.

API mirror of mini_gym/envs/base/curriculum.py (Curriculum :16-68, RewardThresholdCurriculum
:92-124): a 3-D grid of command bins with float64 weights.  The grid, the initial `set_to` and
the per-axis neighbour ranges are computed on the host in float64 with the same expressions as
the reference; the per-event work (success test, weight bump, sampling) runs in the kernels of
csrc/gac.cu, driven by LeggedRobot._resample_commands.
"""
import numpy as np
import torch


class RewardThresholdCurriculum:
    def __init__(self, seed, device="cuda:0", **key_ranges):
        self.rng = np.random.RandomState(seed)  # only used by gac_rng="numpy" replay mode
        self.device = torch.device(device)
        self.cfg = {k: np.linspace(*v) for k, v in key_ranges.items()}           # curriculum.py:27-28
        self.keys = list(key_ranges.keys())
        if len(self.keys) != 3:
            raise NotImplementedError("the device curriculum is 3-D (x_vel, y_vel, yaw_vel)")
        self.bin_sizes = {k: arr[1] - arr[0] for k, arr in self.cfg.items()}      # :30
        self._raw_grid = np.stack(np.meshgrid(*self.cfg.values(), indexing="ij"))  # :32
        self.grid = self._raw_grid.reshape([len(self.keys), -1])                   # :34
        self._l = len(self.grid[0])
        self.ls = {k: len(v) for k, v in self.cfg.items()}
        self.dims = [len(v) for v in self.cfg.values()]
        self.indices = np.arange(self._l)
        # device state
        self.weights_device = torch.zeros(self._l, dtype=torch.float64, device=self.device)
        self.centers_device = torch.from_numpy(np.concatenate(list(self.cfg.values()))).to(self.device)
        self.hit_count = torch.zeros(self._l, dtype=torch.int32, device=self.device)
        self.own_flag = torch.zeros(self._l, dtype=torch.int32, device=self.device)
        self.cdf = torch.zeros(self._l, dtype=torch.float64, device=self.device)
        self._nbr_range = None
        self.nbr_lo = self.nbr_hi = None

    def __len__(self):
        return self._l

    @property
    def weights(self):
        """Host copy (the reference keeps a numpy array; callers only read it for logging)."""
        return self.weights_device.cpu().numpy()

    @weights.setter
    def weights(self, value):
        self.weights_device.copy_(torch.as_tensor(np.asarray(value, dtype=np.float64)))

    def set_to(self, low, high, value=1.0):
        """curriculum.py:17-23."""
        low, high = np.asarray(low, dtype=np.float64), np.asarray(high, dtype=np.float64)
        inds = np.logical_and(self.grid >= low[:, None], self.grid <= high[:, None]).all(axis=0)
        w = self.weights
        w[inds] = value
        self.weights = w

    def neighbour_ranges(self, local_range):
        """Per axis index, the inclusive index range of grid values within +-local_range, evaluated in
        float64 with the comparisons of get_local_bins (curriculum.py:102-108).  Because the grid is
        a tensor product, the neighbourhood of a bin is the product of these three ranges."""
        if self._nbr_range != local_range:
            lo, hi = [], []
            for arr in self.cfg.values():
                for v in arr:
                    ok = np.nonzero(np.logical_and(arr >= v - local_range, arr <= v + local_range))[0]
                    lo.append(ok.min()); hi.append(ok.max())
                    assert len(ok) == ok.max() - ok.min() + 1
            self.nbr_lo = torch.tensor(lo, dtype=torch.int32, device=self.device)
            self.nbr_hi = torch.tensor(hi, dtype=torch.int32, device=self.device)
            self._nbr_range = local_range
        return self.nbr_lo, self.nbr_hi
