"""clodb200: B200-native cluster-LOD DAG builder (host mirror of the reference's builder interface over a C ABI)."""
from .api import ClodLib, ClodbError, Config, load  # noqa: F401
