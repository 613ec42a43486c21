"""Import alias: the sources live in the directory `gcc-nmf_b200/` (named after the reference
repository), which is not a valid Python identifier; this package makes them importable as
`gcc_nmf_b200.<module>` by extending its search path."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'gcc-nmf_b200'))
