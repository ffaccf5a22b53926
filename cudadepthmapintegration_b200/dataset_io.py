"""VTK-free writers for the reference's on-disk inputs (tests, examples): .vti depth maps with the point
arrays "Depths" / "Best Cost Values" / "Color" (Sources/ReconstructionData.cxx:95,144,146), .krtd camera
files (3 lines K, blank, 3 lines R, blank, 1 line T -- Sources/Helper.h:105-168) and the list files
(one file name per line, resolved against the list's directory -- Helper.h:60-100)."""
from __future__ import annotations

import base64
import os
import struct
import zlib

import numpy as np


def _block(raw: bytes, hfmt: str, compress: bool, block_size: int = 32768) -> tuple[bytes, bytes]:
    """(header, payload) of one VTK XML binary data block: uncompressed = byte count + data; compressed = (number of
    blocks, block size, size of the last block, compressed sizes) + the zlib streams."""
    if not compress:
        return struct.pack(hfmt, len(raw)), raw
    chunks = [raw[o:o + block_size] for o in range(0, len(raw), block_size)] or [b""]
    comp = [zlib.compress(c) for c in chunks]
    last = len(chunks[-1]) if len(chunks[-1]) != block_size else 0
    head = struct.pack(hfmt, len(chunks)) + struct.pack(hfmt, block_size) + struct.pack(hfmt, last)
    head += b"".join(struct.pack(hfmt, len(c)) for c in comp)
    return head, b"".join(comp)


def write_vti(path, depths, best_cost=None, color=None, ascii=False, encoding=None, compress=False, header_type="UInt32",
              header_with_data=False):
    """depths / best_cost: (H, W) float64, bottom-up rows; color: (H, W, 3) uint8.
    encoding: "raw" (appended raw, the default), "ascii", "base64" (appended base64, what VTK's writers produce by
    default, with compress=True), "binary" (inline base64).  compress: zlib blocks (vtkZLibDataCompressor).
    header_with_data: base64-encode an uncompressed block's byte count together with its data instead of on its
    own (both layouts occur in the wild)."""
    encoding = encoding or ("ascii" if ascii else "raw")
    hfmt = "<I" if header_type == "UInt32" else "<Q"
    H, W = depths.shape
    arrays = [("Depths", "Float64", 1, np.ascontiguousarray(depths, dtype=np.float64))]
    if best_cost is not None:
        arrays.append(("Best Cost Values", "Float64", 1, np.ascontiguousarray(best_cost, dtype=np.float64)))
    if color is not None:
        arrays.append(("Color", "UInt8", 3, np.ascontiguousarray(color, dtype=np.uint8)))
    comp_attr = ' compressor="vtkZLibDataCompressor"' if compress and encoding != "ascii" else ""
    head = ['<?xml version="1.0"?>',
            f'<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian" header_type="{header_type}"{comp_attr}>',
            f'  <ImageData WholeExtent="0 {W - 1} 0 {H - 1} 0 0" Origin="0 0 0" Spacing="1 1 1">',
            f'    <Piece Extent="0 {W - 1} 0 {H - 1} 0 0">', '      <PointData>']

    def b64(block_head: bytes, payload: bytes) -> bytes:
        if header_with_data and not compress:
            return base64.b64encode(block_head + payload)
        return base64.b64encode(block_head) + base64.b64encode(payload)

    blob = b""
    for name, typ, comps, a in arrays:
        nc = f' NumberOfComponents="{comps}"' if comps > 1 else ""
        if encoding == "ascii":
            txt = " ".join(repr(float(x)) if typ != "UInt8" else str(int(x)) for x in a.reshape(-1))
            head.append(f'        <DataArray type="{typ}" Name="{name}"{nc} format="ascii">{txt}</DataArray>')
            continue
        bh, payload = _block(a.tobytes(), hfmt, compress)
        if encoding == "binary":
            head.append(f'        <DataArray type="{typ}" Name="{name}"{nc} format="binary">\n          '
                        + b64(bh, payload).decode() + '\n        </DataArray>')
        else:
            head.append(f'        <DataArray type="{typ}" Name="{name}"{nc} format="appended" offset="{len(blob)}"/>')
            blob += (bh + payload) if encoding == "raw" else b64(bh, payload)
    head += ['      </PointData>', '    </Piece>', '  </ImageData>']
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode())
        if encoding in ("raw", "base64"):
            f.write(f'  <AppendedData encoding="{encoding}">\n   _'.encode() + blob + b'\n  </AppendedData>\n')
        f.write(b"</VTKFile>\n")


def write_krtd(path, K4, RT4):
    K = np.asarray(K4, dtype=np.float64).reshape(4, 4)
    RT = np.asarray(RT4, dtype=np.float64).reshape(4, 4)
    with open(path, "w") as f:
        for i in range(3):
            f.write(" ".join(repr(float(K[i, j])) for j in range(3)) + "\n")
        f.write("\n")
        for i in range(3):
            f.write(" ".join(repr(float(RT[i, j])) for j in range(3)) + "\n")
        f.write("\n")
        f.write(" ".join(repr(float(RT[i, 3])) for i in range(3)) + "\n")
        f.write("\n0\n")          # distortion line of the kwiver .krtd format (ignored by the reference)


def write_dataset(folder, depths, best_cost, colors, K, RT, vti_list="vtiList.txt", krtd_list="kList.txt", ascii_views=(),
                  vti_options=None):
    """One .vti + one .krtd per view and the two list files, in `folder`.  vti_options: per-view dicts of extra
    write_vti arguments (encoding, compress, header_type, header_with_data), cycled over the views."""
    os.makedirs(folder, exist_ok=True)
    n = len(depths)
    with open(os.path.join(folder, vti_list), "w") as fv, open(os.path.join(folder, krtd_list), "w") as fk:
        for v in range(n):
            vname, kname = f"view_{v:04d}.vti", f"view_{v:04d}.krtd"
            extra = dict(vti_options[v % len(vti_options)]) if vti_options else {}
            write_vti(os.path.join(folder, vname), depths[v], None if best_cost is None else best_cost[v],
                      None if colors is None else colors[v], ascii=v in ascii_views, **extra)
            write_krtd(os.path.join(folder, kname), K[v], RT[v])
            # the reference takes the LAST space-separated token of each line (Helper.h:86-97)
            fv.write(f"{v} {vname}\n")
            fk.write(f"{kname}\n")
        fv.write("\n")
