"""Build a tuning variant:  python tools/build_variant.py NAME -DFLAG ...  ->  csrc/libfutbol_b200_NAME.so (select with FUTBOL_B200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gym_futbol_b200.build import build_extension
name=sys.argv[1]; flags=tuple(sys.argv[2:])
print(build_extension(extra_flags=flags, out="libfutbol_b200_%s.so"%name))
