// coop.cuh placeholder
