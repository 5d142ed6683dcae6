from .optflow_registrator import OptFlowRegistrator
from .warper import Warper
