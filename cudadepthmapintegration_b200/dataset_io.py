"""VTK-free writers AND readers for the reference's on-disk inputs: .vti depth maps with the point
arrays "Depths" / "Best Cost Values" / "Color" (Sources/ReconstructionData.cxx:95,144,146), .krtd camera
files (3 lines K, blank, 3 lines R, blank, 1 line T -- Sources/Helper.h:105-168) and the list files
(one file name per line, resolved against the list's directory -- Helper.h:60-100)."""
from __future__ import annotations

import base64
import os
import re
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


# ---- readers (the Python mirrors of the two operators take file lists like the reference's filters) ----------

def extract_all_file_path(list_file: str) -> list[str]:
    """help::ExtractAllFilePath (Sources/Helper.h:60-100): one file per line = the line's last blank-separated token
    (with the reference's tokenisation: one trailing blank does not open a token, two do), relative to the list's
    directory, or to the working directory when the list was named without one."""
    try:
        with open(list_file, "r", newline="") as f:
            text = f.read()
    except OSError:
        import sys
        print(f"Unable to open : {list_file}", file=sys.stderr)
        return []
    norm = list_file.replace("\\", "/")
    cut = norm.rfind("/")
    if cut < 0:
        directory = os.getcwd()
    elif cut == 0:
        directory = "/"
    elif cut == 2 and norm[1] == ":":
        directory = norm[:2] + "/"
    else:
        directory = norm[:cut]
    out = []
    for line in text.split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if not line:
            continue
        tokens = line.split(" ")
        if len(tokens) > 1 and tokens[-1] == "" :
            tokens.pop()                       # std::getline does not produce a token after one trailing blank
        out.append(directory + "/" + tokens[-1])
    return out


def read_krtd(path: str):
    """help::ReadKrtdFile (Helper.h:105-168) -> (K4, RT4) as row-major float64[16]: K on lines 1-3, line 4 skipped,
    R on lines 5-7, line 8 skipped, T on line 9; three numbers per line, a missing one reads as 0."""
    with open(path, "r") as f:
        lines = f.read().split("\n")
    lines += [""] * 9

    def three(line):
        vals = []
        for tok in line.split():
            try:
                vals.append(float(tok))
            except ValueError:
                break
        return (vals + [0.0, 0.0, 0.0])[:3]
    K = np.eye(4)
    RT = np.zeros((4, 4))
    for i in range(3):
        K[i, :3] = three(lines[i])
        RT[i, :3] = three(lines[4 + i])
    RT[:3, 3] = three(lines[8])
    RT[3, 3] = 1.0
    return K.reshape(16), RT.reshape(16)


def _attr(tag: str, name: str) -> str:
    m = re.search(r'(?:^|\s)' + re.escape(name) + r'="([^"]*)"', tag)
    return m.group(1) if m else ""


def _decode_block(buf: bytes, is_base64: bool, compressed: bool, hfmt: str, want: int, what: str) -> bytes:
    hs = struct.calcsize(hfmt)

    def b64chars(n):
        return (n + 2) // 3 * 4

    def b64(data: bytes, nchars: int) -> bytes:
        chunk = data[:nchars]
        m = re.match(rb"[A-Za-z0-9+/]*", chunk)
        body = m.group(0)
        body = body[:len(body) // 4 * 4] + (body[len(body) // 4 * 4:] + b"===")[:4] if len(body) % 4 else body
        return base64.b64decode(body)
    if not compressed:
        if not is_base64:
            n = struct.unpack_from(hfmt, buf, 0)[0]
            if n != want or len(buf) < hs + n:
                raise ValueError(f"{what}: byte count does not match the image extent")
            return buf[hs:hs + n]
        head_chars = buf[:b64chars(hs)]
        separate = b"=" in head_chars
        n = struct.unpack_from(hfmt, b64(buf, b64chars(hs)), 0)[0]
        if n != want:
            raise ValueError(f"{what}: byte count does not match the image extent")
        data = b64(buf[b64chars(hs):], b64chars(want)) if separate else b64(buf, b64chars(hs + want))[hs:]
        if len(data) < want:
            raise ValueError(f"{what}: truncated block")
        return data[:want]
    if is_base64:
        nb = struct.unpack_from(hfmt, b64(buf, b64chars(3 * hs)), 0)[0]
        hchars = b64chars((3 + nb) * hs)
        head = b64(buf, hchars)
        sizes = [struct.unpack_from(hfmt, head, (3 + b) * hs)[0] for b in range(nb)]
        data = b64(buf[hchars:], b64chars(sum(sizes)))
    else:
        nb = struct.unpack_from(hfmt, buf, 0)[0]
        head = buf[:(3 + nb) * hs]
        sizes = [struct.unpack_from(hfmt, head, (3 + b) * hs)[0] for b in range(nb)]
        data = buf[(3 + nb) * hs:]
    block, last = struct.unpack_from(hfmt, head, hs)[0], struct.unpack_from(hfmt, head, 2 * hs)[0]
    out, pos = [], 0
    for b, cs in enumerate(sizes):
        chunk = zlib.decompress(data[pos:pos + cs])
        if len(chunk) != (last if (b == nb - 1 and last) else block):
            raise ValueError(f"{what}: corrupt compressed block")
        out.append(chunk)
        pos += cs
    raw = b"".join(out)
    if len(raw) != want:
        raise ValueError(f"{what}: uncompressed size does not match the image extent")
    return raw


def read_vti(path: str):
    """(depths (H, W) f64, best_cost (H, W) f64 or None, color (H, W, 3) u8 or None) of a depth-map .vti, bottom-up
    rows: the point arrays ReconstructionData reads (Sources/ReconstructionData.cxx:95,144,146).  Same layouts as the
    C++ reader (csrc/host/DmiVti.h)."""
    with open(path, "rb") as f:
        s = f.read()
    vf = s.find(b"<VTKFile")
    if vf < 0:
        raise ValueError(f"{path}: not a VTK XML file")
    vtag = s[vf:s.index(b">", vf)].decode("latin-1")
    if _attr(vtag, "type") != "ImageData":
        raise ValueError(f"{path}: VTKFile type is not ImageData")
    if _attr(vtag, "byte_order") not in ("", "LittleEndian"):
        raise ValueError(f"{path}: big-endian .vti not supported")
    comp = _attr(vtag, "compressor")
    if comp not in ("", "vtkZLibDataCompressor"):
        raise ValueError(f"{path}: compressor {comp} is not supported (zlib only)")
    hfmt = "<Q" if _attr(vtag, "header_type") == "UInt64" else "<I"
    it = s.find(b"<ImageData")
    itag = s[it:s.index(b">", it)].decode("latin-1")
    e = [int(x) for x in _attr(itag, "WholeExtent").split()]
    W, H = e[1] - e[0] + 1, e[3] - e[2] + 1
    if W <= 0 or H <= 0 or e[5] != e[4]:
        raise ValueError(f"{path}: expected a 2-D image extent")
    app, app_b64 = None, False
    ad = s.find(b"<AppendedData")
    if ad >= 0:
        atag = s[ad:s.index(b">", ad)].decode("latin-1")
        us = s.find(b"_", s.index(b">", ad))
        enc = _attr(atag, "encoding")
        if us >= 0 and enc in ("", "base64", "raw"):
            app, app_b64 = us + 1, enc != "raw"
    pd0, pd1 = s.find(b"<PointData"), s.find(b"</PointData>")
    if pd0 < 0 or pd1 < 0:
        raise ValueError(f"{path}: no <PointData>")
    found = {}
    pos = pd0
    while True:
        pos = s.find(b"<DataArray", pos)
        if pos < 0 or pos > pd1:
            break
        te = s.index(b">", pos)
        tag = s[pos:te].decode("latin-1")
        name, typ, fmt = _attr(tag, "Name"), _attr(tag, "type"), _attr(tag, "format")
        comps = int(_attr(tag, "NumberOfComponents") or 1)
        pos = te
        if name not in ("Depths", "Best Cost Values", "Color"):
            continue
        dt = {"Float64": np.float64, "Float32": np.float32, "UInt8": np.uint8}.get(typ)
        if dt is None:
            raise ValueError(f"{path}: array '{name}' has an unsupported type")
        count = W * H * comps
        if fmt == "ascii":
            txt = s[te + 1:s.index(b"</DataArray>", te)].split()
            arr = np.array([float(x) for x in txt[:count]]).astype(dt)
        elif fmt in ("appended", "binary"):
            if fmt == "appended":
                if app is None:
                    raise ValueError(f"{path}: appended data section missing or in an unknown encoding")
                buf, is_b64 = s[app + int(_attr(tag, "offset") or 0):], app_b64
            else:
                buf, is_b64 = s[te + 1:].lstrip(), True
            arr = np.frombuffer(_decode_block(buf, is_b64, bool(comp), hfmt, count * np.dtype(dt).itemsize, f"{path}: array '{name}'"), dtype=dt)
        else:
            raise ValueError(f"{path}: format '{fmt}' is not supported")
        if arr.size != count:
            raise ValueError(f"{path}: array '{name}' has the wrong size")
        found[name] = arr
    if "Depths" not in found:
        raise ValueError(f"{path}: no 'Depths' array")
    depths = found["Depths"].astype(np.float64).reshape(H, W)
    cost = found["Best Cost Values"].astype(np.float64).reshape(H, W) if "Best Cost Values" in found else None
    color = None
    if "Color" in found:
        if found["Color"].dtype != np.uint8 or found["Color"].size != W * H * 3:
            raise ValueError(f"{path}: Color must be UInt8 x 3")
        color = found["Color"].reshape(H, W, 3).copy()
    return depths, cost, color


def load_dataset(vti_list: str, krtd_list: str, need_color: bool = False):
    """Everything the two operators read from disk, in list order: (depths [n,H,W] f64, best_cost [n,H,W] f64 or
    None, colors [n,H,W,3] u8 or None, K [n,16], RT [n,16]).  Like the reference, fewer .krtd than .vti entries is
    an error (vtkCudaReconstructionFilter / CudaReconstruction.cu:308-312, MeshColoration.cxx:59-63); every depth
    map must have the size of the first one; best-cost maps are used only when every view has one."""
    vti, krtd = extract_all_file_path(vti_list), extract_all_file_path(krtd_list)
    if len(vti) == 0 or len(krtd) < len(vti):
        raise ValueError("There is no enough vti files, please check your vtiList.txt and krtdList.txt")
    depths, costs, colors, Ks, RTs = [], [], [], [], []
    for v, path in enumerate(vti):
        d, c, col = read_vti(path)
        if depths and d.shape != depths[0].shape:
            raise ValueError(f"{path}: depth map size differs from the first one")
        if need_color and col is None:
            raise ValueError(f"{path}: no 'Color' array exists")
        depths.append(d); costs.append(c); colors.append(col)
        K, RT = read_krtd(krtd[v])
        Ks.append(K); RTs.append(RT)
    have_cost = all(c is not None for c in costs)
    have_color = all(c is not None for c in colors)
    return (np.stack(depths), np.stack(costs) if have_cost else None, np.stack(colors) if have_color else None,
            np.stack(Ks), np.stack(RTs))
