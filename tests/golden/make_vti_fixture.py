"""Writes tests/golden/vtk_style_view.vti + .krtd + list files: ONE view in the layout vtkXMLImageDataWriter produces with
its defaults in VTK 6-8 (file version 0.1, LittleEndian, UInt32 block headers, vtkZLibDataCompressor, appended data in base64
with the block header and the compressed blocks encoded as two separate base64 streams, `offset` = characters after the `_`,
RangeMin / RangeMax attributes, an empty CellData element), assembled here from the VTK file-format description with
struct / zlib / base64 only -- NOT with cudadepthmapintegration_b200.dataset_io, so that the readers under test and the
writer they are usually round-tripped against do not share an author's reading of the format.  The arrays are seeded;
tests/test_host_cli.py regenerates them and compares what the readers return.

    python tests/golden/make_vti_fixture.py
"""
import base64
import os
import struct
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
W, H, BLOCK = 37, 23, 4096            # 37*23*8 = 6808 bytes: two blocks, the second one ragged


def arrays():
    rng = np.random.RandomState(20261017)
    depths = rng.uniform(0.5, 4.0, size=(H, W))
    depths[rng.uniform(size=(H, W)) < 0.2] = -1.0
    cost = rng.uniform(0.0, 0.2, size=(H, W))
    color = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    K = np.array([[31.5, 0.25, 18.0], [0.0, 30.75, 11.5], [0.0, 0.0, 1.0]])
    R = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])
    T = np.array([0.125, -0.5, 3.0])
    return depths, cost, color, K, R, T


def encode_array(raw: bytes) -> str:
    blocks = [raw[o:o + BLOCK] for o in range(0, len(raw), BLOCK)]
    comp = [zlib.compress(b) for b in blocks]
    last = len(raw) % BLOCK                      # 0 when the last block is full (VTK's convention)
    header = struct.pack("<%dI" % (3 + len(comp)), len(blocks), BLOCK, last, *[len(c) for c in comp])
    return base64.b64encode(header).decode() + base64.b64encode(b"".join(comp)).decode()


def main():
    depths, cost, color, K, R, T = arrays()
    parts, offsets, pos = [], [], 0
    for a in (depths.astype("<f8"), cost.astype("<f8"), color):
        enc = encode_array(a.tobytes())
        offsets.append(pos)
        parts.append(enc)
        pos += len(enc)
    ext = f"0 {W - 1} 0 {H - 1} 0 0"
    xml = (
        '<?xml version="1.0"?>\n'
        '<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian" compressor="vtkZLibDataCompressor">\n'
        f'  <ImageData WholeExtent="{ext}" Origin="0 0 0" Spacing="1 1 1">\n'
        f'  <Piece Extent="{ext}">\n'
        '    <PointData Scalars="Depths">\n'
        f'      <DataArray type="Float64" Name="Depths" format="appended" RangeMin="{depths.min()}" RangeMax="{depths.max()}" offset="{offsets[0]}"/>\n'
        f'      <DataArray type="Float64" Name="Best Cost Values" format="appended" RangeMin="{cost.min()}" RangeMax="{cost.max()}" offset="{offsets[1]}"/>\n'
        f'      <DataArray type="UInt8" Name="Color" NumberOfComponents="3" format="appended" RangeMin="0" RangeMax="441.67" offset="{offsets[2]}"/>\n'
        '    </PointData>\n'
        '    <CellData>\n'
        '    </CellData>\n'
        '  </Piece>\n'
        '  </ImageData>\n'
        '  <AppendedData encoding="base64">\n'
        '   _' + "".join(parts) + '\n'
        '  </AppendedData>\n'
        '</VTKFile>\n')
    open(os.path.join(HERE, "vtk_style_view.vti"), "w").write(xml)
    rows = lambda m: "\n".join(" ".join(repr(float(x)) for x in r) for r in m)
    open(os.path.join(HERE, "vtk_style_view.krtd"), "w").write(rows(K) + "\n\n" + rows(R) + "\n\n" + " ".join(repr(float(x)) for x in T) + "\n\n0\n")
    open(os.path.join(HERE, "vtk_style_vtiList.txt"), "w").write("0 vtk_style_view.vti\n")
    open(os.path.join(HERE, "vtk_style_kList.txt"), "w").write("vtk_style_view.krtd\n")


if __name__ == "__main__":
    main()
