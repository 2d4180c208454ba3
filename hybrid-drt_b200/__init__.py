"""hybdrt_b200 -- B200-native batched DRT/DOP inversion engine behind hybrid-drt's Python API."""
__version__ = '0.1.0'
