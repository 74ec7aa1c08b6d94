"""gymapi stand-in: asset facts from the URDF, simulator state as torch tensors."""
import xml.etree.ElementTree as ET
import os

import numpy as np
import torch

SIM_PHYSX = 1
DOMAIN_SIM = 2
KEY_ESCAPE = 0
KEY_V = 1
IMAGE_COLOR = 0


class Vec3:
    def __init__(self, x=0., y=0., z=0.):
        self.x, self.y, self.z = float(x), float(y), float(z)


class Quat:
    def __init__(self, x=0., y=0., z=0., w=1.):
        self.x, self.y, self.z, self.w = x, y, z, w


class Transform:
    def __init__(self, p=None, r=None):
        self.p = p or Vec3()
        self.r = r or Quat()


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class SimParams:
    """`dt` is a C float in the real binding: 0.005 reads back as 0.004999999888..."""

    def __init__(self):
        self._dt = float(np.float32(1.0 / 60.0))
        self.substeps = 2
        self.use_gpu_pipeline = False
        self.physx = _Bag()

    @property
    def dt(self):
        return self._dt

    @dt.setter
    def dt(self, v):
        self._dt = float(np.float32(v))


class PlaneParams(_Bag):
    pass


class HeightFieldParams(_Bag):
    def __init__(self):
        super().__init__(transform=Transform())


class TriangleMeshParams(_Bag):
    def __init__(self):
        super().__init__(transform=Transform())


class AssetOptions(_Bag):
    pass


class CameraProperties(_Bag):
    pass


def _parse_urdf(path, collapse_fixed):
    """Link / DOF tables after fixed-joint collapsing (depth-first, URDF order)."""
    root = ET.parse(path).getroot()
    joints = root.findall('joint')
    children = {}
    child_links = set()
    for j in joints:
        children.setdefault(j.find('parent').get('link'), []).append(j)
        child_links.add(j.find('child').get('link'))
    links = [l.get('name') for l in root.findall('link')]
    base = [l for l in links if l not in child_links][0]
    bodies, dofs = [], []

    def visit(link, keep):
        if keep:
            bodies.append(link)
        for j in children.get(link, []):
            jtype = j.get('type')
            child = j.find('child').get('link')
            if jtype == 'fixed':
                kept = (not collapse_fixed) or j.get('dont_collapse') == 'true'
                visit(child, kept)
            else:
                lim = j.find('limit')
                dofs.append(dict(name=j.get('name'), lower=float(lim.get('lower')), upper=float(lim.get('upper')),
                                 velocity=float(lim.get('velocity')), effort=float(lim.get('effort'))))
                visit(child, True)

    visit(base, True)
    return bodies, dofs


class _Asset:
    def __init__(self, path, options):
        self.bodies, self.dofs = _parse_urdf(path, getattr(options, 'collapse_fixed_joints', True))


class _Sim:
    def __init__(self, device):
        self.device = device
        self.envs = []
        self.asset = None
        self.tensors = {}


class _Env:
    def __init__(self, idx):
        self.idx = idx


class Gym:
    """Every method the reference calls with a meaningful return is defined;
    anything else (simulate, refresh_*, set_*, viewer calls...) is a no-op."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)

        def _noop(*a, **k):
            return None
        return _noop

    def create_sim(self, compute_device, graphics_device, physics_engine, sim_params):
        from . import gymutil
        on_gpu = gymutil.LAST_DEVICE_TYPE == 'cuda' and sim_params.use_gpu_pipeline
        dev = 'cuda:%d' % compute_device if on_gpu else 'cpu'
        self._sim = _Sim(dev)
        return self._sim

    def load_asset(self, sim, root, file, options):
        sim.asset = _Asset(os.path.join(root, file), options)
        return sim.asset

    def get_asset_dof_count(self, asset):
        return len(asset.dofs)

    def get_asset_rigid_body_count(self, asset):
        return len(asset.bodies)

    def get_asset_dof_properties(self, asset):
        props = np.zeros(len(asset.dofs), dtype=[('lower', 'f4'), ('upper', 'f4'), ('velocity', 'f4'),
                                                 ('effort', 'f4')])
        for i, d in enumerate(asset.dofs):
            props[i] = (d['lower'], d['upper'], d['velocity'], d['effort'])
        return props

    def get_asset_rigid_shape_properties(self, asset):
        return [_Bag(friction=1.0, restitution=0.0) for _ in range(max(2, len(asset.bodies)))]

    def get_asset_rigid_body_names(self, asset):
        return list(asset.bodies)

    def get_asset_dof_names(self, asset):
        return [d['name'] for d in asset.dofs]

    def create_env(self, sim, lower, upper, per_row):
        env = _Env(len(sim.envs))
        sim.envs.append(env)
        return env

    def create_actor(self, env, asset, pose, name, group, filt, seg=0):
        return 0

    def get_actor_rigid_body_properties(self, env, actor):
        return [_Bag(mass=1.0, com=Vec3()) for _ in self._sim.asset.bodies]

    def find_actor_index(self, env, name, domain):
        return env.idx

    def get_actor_dof_index(self, env, actor, dof_idx, domain):
        return env.idx * len(self._sim.asset.dofs) + dof_idx

    def find_actor_rigid_body_index(self, env, actor, name, domain):
        return env.idx * len(self._sim.asset.bodies) + self._sim.asset.bodies.index(name)

    def find_actor_rigid_body_handle(self, env, actor, name):
        return self._sim.asset.bodies.index(name)

    def _tensor(self, sim, key, rows, cols):
        if key not in sim.tensors:
            sim.tensors[key] = torch.zeros(rows, cols, dtype=torch.float, device=sim.device)
        return sim.tensors[key]

    def acquire_actor_root_state_tensor(self, sim):
        t = self._tensor(sim, 'root', len(sim.envs), 13)
        if not t.any():
            t[:, 6] = 1.0
        return t

    def acquire_dof_state_tensor(self, sim):
        return self._tensor(sim, 'dof', len(sim.envs) * len(sim.asset.dofs), 2)

    def acquire_net_contact_force_tensor(self, sim):
        return self._tensor(sim, 'contact', len(sim.envs) * len(sim.asset.bodies), 3)

    def acquire_rigid_body_state_tensor(self, sim):
        return self._tensor(sim, 'rb', len(sim.envs) * len(sim.asset.bodies), 13)


def acquire_gym():
    return Gym()
