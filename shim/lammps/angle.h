// forwarding header: in a real LAMMPS tree the upstream "angle.h" is found instead (see lammps_shim.h)
#include "lammps_shim.h"
