"""gymtorch stand-in: tensor handles ARE torch tensors."""


def wrap_tensor(handle):
    return handle


def unwrap_tensor(t):
    return t
