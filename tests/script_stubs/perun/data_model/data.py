import enum


class MetricType(enum.Enum):
    ENERGY = "energy"
    RUNTIME = "runtime"
    CO2 = "co2"
    MONEY = "money"
    GPU_POWER = "gpu_power"
    GPU_UTIL = "gpu_util"
    GPU_MEM = "gpu_mem"


class DataNode:
    def __init__(self):
        self.metrics = {}
        self.nodes = {}
        self.raw_data = None
