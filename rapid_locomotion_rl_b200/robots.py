"""Robot constants the hot path needs (body / DOF tables and URDF limits).

The reference obtains these from Isaac Gym at env construction
(mini_gym/envs/base/legged_robot.py:1190-1207, 1283-1300, 501-515).  Isaac Gym is a
closed simulator that is not part of this build, so the two robots the reference ships are
described here as data, taken from its URDFs:
  resources/robots/mini_cheetah/urdf/mini_cheetah.urdf:103,133,162 (limits), 13 bodies after
  fixed-joint collapse; resources/robots/go1/urdf/go1.urdf:96,138,166 (limits), 17 bodies (feet
  kept by dont_collapse, :188).
"""
from dataclasses import dataclass, field
from typing import List

LEGS = ("FR", "FL", "RR", "RL")


@dataclass
class RobotSpec:
    name: str
    body_names: List[str]
    dof_names: List[str]
    dof_lower: List[float]
    dof_upper: List[float]
    dof_velocity: List[float]
    dof_effort: List[float]
    default_friction: float = 1.0
    default_restitution: float = 0.0
    default_body_mass: float = 1.0

    @property
    def num_bodies(self):
        return len(self.body_names)

    @property
    def num_dof(self):
        return len(self.dof_names)

    def bodies_matching(self, substrings):
        """Indices of bodies whose name contains any substring, grouped in substring order
        (the order legged_robot.py:1201-1207 builds its name lists in)."""
        if isinstance(substrings, str):
            substrings = [substrings]
        out = []
        for sub in substrings:
            out.extend(i for i, n in enumerate(self.body_names) if sub in n)
        return out


def _legged(name, parts, hip, thigh, calf):
    bodies = ["base"]
    dofs, lo, hi, vel, eff = [], [], [], [], []
    for leg in LEGS:
        bodies.extend("%s_%s" % (leg, p) for p in parts)
        for joint, lim in (("hip", hip), ("thigh", thigh), ("calf", calf)):
            dofs.append("%s_%s_joint" % (leg, joint))
            lo.append(lim[0]); hi.append(lim[1]); vel.append(lim[2]); eff.append(lim[3])
    return RobotSpec(name, bodies, dofs, lo, hi, vel, eff)


# (lower, upper, velocity, effort)
MINI_CHEETAH = _legged("mini_cheetah", ("hip", "thigh", "calf"),
                       hip=(-1.6, 1.6, 40.0, 18.0), thigh=(-2.6, 2.6, 40.0, 18.0), calf=(-2.6, 2.6, 26.0, 26.0))
GO1 = _legged("go1", ("hip", "thigh", "calf", "foot"),
              hip=(-0.802851455917, 0.802851455917, 50.0, 33.5),
              thigh=(-1.0471975512, 4.18879020479, 28.0, 33.5),
              calf=(-2.69653369433, -0.916297857297, 28.0, 33.5))

ROBOTS = {"mini_cheetah": MINI_CHEETAH, "go1": GO1}


def robot_for_asset(asset_file: str) -> RobotSpec:
    for key, spec in ROBOTS.items():
        if key in asset_file:
            return spec
    raise ValueError("no robot table for asset %r (known: %s)" % (asset_file, ", ".join(ROBOTS)))
