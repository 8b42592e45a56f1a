"""Wraps a metadata blob written by clodb200_artifactsSaveCache into the .usdc stage the renderer opens
(BasicRenderer/src/Import/CLodCache.cpp:528-574: prim /CLodCache with clodSchemaVersion, clodBuildConfigHash and
`uchar[] clodBlob`, plus one Scope /CLodCache/Groups/g_<i> with `uint groupIndex` per group).

Needs OpenUSD's Python bindings (`pxr`), which this image does not have: the script is NOT exercised by the test suite
("parity unpinned"). Run it on a machine with OpenUSD 25.05:

    python tools/wrap_usdc.py <metadata blob> <out.usdc>
"""
import struct
import sys


def main(blob_path: str, out_path: str) -> None:
    from pxr import Sdf, Usd, Vt  # noqa: imported here so that the module can be read without OpenUSD

    blob = open(blob_path, "rb").read()
    schema, config_hash, group_count = struct.unpack_from("<IQQ", blob, 0)  # SerializeMetadata, CLodCache.cpp:169-176
    stage = Usd.Stage.CreateNew(out_path, Usd.Stage.LoadNone)
    root = stage.DefinePrim(Sdf.Path("/CLodCache"), "Scope")
    stage.DefinePrim(Sdf.Path("/CLodCache/Groups"), "Scope")
    root.CreateAttribute("clodSchemaVersion", Sdf.ValueTypeNames.Int, True).Set(int(schema))
    root.CreateAttribute("clodBuildConfigHash", Sdf.ValueTypeNames.Int64, True).Set(config_hash - (1 << 64) if config_hash >> 63 else config_hash)
    root.CreateAttribute("clodBlob", Sdf.ValueTypeNames.UCharArray, True).Set(Vt.UCharArray(list(blob)))
    for g in range(group_count):
        prim = stage.DefinePrim(Sdf.Path(f"/CLodCache/Groups/g_{g}"), "Scope")
        prim.CreateAttribute("groupIndex", Sdf.ValueTypeNames.UInt, True).Set(g)
    stage.GetRootLayer().Save()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
