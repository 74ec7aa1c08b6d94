"""gymutil stand-in."""


LAST_DEVICE_TYPE = 'cpu'


def parse_device_str(device_str):
    global LAST_DEVICE_TYPE
    device_str = str(device_str)
    LAST_DEVICE_TYPE = device_str.split(':')[0]
    if device_str == 'cpu' or device_str == 'cuda':
        return device_str, 0
    kind, idx = device_str.split(':')
    return kind, int(idx)


def parse_sim_config(sim_cfg, sim_params):
    if 'dt' in sim_cfg:
        sim_params.dt = sim_cfg['dt']
    if 'use_gpu_pipeline' in sim_cfg:
        sim_params.use_gpu_pipeline = sim_cfg['use_gpu_pipeline']
    if 'substeps' in sim_cfg:
        sim_params.substeps = sim_cfg['substeps']


class WireframeSphereGeometry:
    def __init__(self, *a, **k):
        pass


def draw_lines(*a, **k):
    pass
