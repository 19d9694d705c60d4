from .disp_imop import DispImOpeartion
from .int_energy_exop import IntEnergyExOperation
from .volume_exop import VolumeExOperation
