"""Stages the UNMODIFIED reference into oracle/_ref/ (git-ignored; it travels to the GPU box with gpurun).

TEST / BASELINE INFRASTRUCTURE ONLY - nothing under rapid_locomotion_rl_b200/ may import it.

    python oracle/make_ref.py            # /root/reference/{mini_gym,mini_gym_learn} -> oracle/_ref/

The reference is pure Python (setup.py:1-17: packages mini_gym, mini_gym_learn), so "building" it is placing the two
package trees on an import path.  (`pip install --target` of the source drops mini_gym/envs/world - the directory has
no __init__.py and setup.py uses find_packages() - so the trees are copied as they are.)  Nothing is patched: `bench.py --impl reference` and the `cpu_baseline` /
`ppo.reference` legs import it through the fake-simulator shims in tests/_ref_shims (our own code: a ~90-line
`isaacgym` stand-in plus stubs for gym / params_proto / ml_logger / matplotlib), exactly like
tests/golden/make_golden.py does in the build container.  /root/reference does not exist on the GPU box; this
staged copy is what lets the reference arm be the reference there (`cpu_baseline.kind = "reference"`).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("RL_REFERENCE_ROOT", "/root/reference")
PACKAGES = ("mini_gym", "mini_gym_learn")
CHECKPOINT = "runs/rapid-locomotion/example/train/201852.132488/checkpoints/ac_weights_last.pt"


def staged():
    return all(os.path.isdir(os.path.join(DEST, p)) for p in PACKAGES + ("resources",)) and \
        os.path.isfile(os.path.join(DEST, CHECKPOINT))


def make(force=False):
    """Returns the staged path, or None when the reference source is not on this machine."""
    if staged() and not force:
        return DEST
    if not os.path.isdir(os.path.join(SRC, "mini_gym")):
        return DEST if staged() else None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    for p in PACKAGES:
        shutil.copytree(os.path.join(SRC, p), os.path.join(DEST, p), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    # robot descriptions: the fake simulator reads link / joint names from the .urdf files (meshes are not needed)
    for dirpath, _, files in os.walk(os.path.join(SRC, "resources", "robots")):
        for f in files:
            if f.endswith(".urdf"):
                out = os.path.join(DEST, os.path.relpath(dirpath, SRC))
                os.makedirs(out, exist_ok=True)
                shutil.copy(os.path.join(dirpath, f), out)
    # the trained example policy the reference ships (runs/.../checkpoints/ac_weights_last.pt, 2.4 MB of weights): the
    # checkpoint-compatibility test loads it into the product ActorCritic
    ck = os.path.join(SRC, CHECKPOINT)
    if os.path.isfile(ck):
        os.makedirs(os.path.dirname(os.path.join(DEST, CHECKPOINT)), exist_ok=True)
        shutil.copy(ck, os.path.join(DEST, CHECKPOINT))
    return DEST


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
