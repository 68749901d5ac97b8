"""B200-native VEGAS Monte Carlo integration (API of N3PDF/vegasflow v1.4.0)."""

from vegasflow_b200.configflow import DTYPE, DTYPEINT, float_me, int_me, run_eager
from vegasflow_b200 import integrands
from vegasflow_b200.plain import PlainFlow, plain_sampler, plain_wrapper
from vegasflow_b200.vflow import VegasFlow, vegas_sampler, vegas_wrapper
from vegasflow_b200.vflowplus import VegasFlowPlus, vegasflowplus_sampler, vegasflowplus_wrapper

__version__ = "0.1.0"
