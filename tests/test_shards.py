"""Raw fp32 shard format + dependency-free .vtu reader (desmo_b200/shards.py): the reference's read_velocity_data (CYL:39-85) as a
one-off conversion, per-rank slabs of mesh points, and the device pre-processing + POD fed from them."""
import base64
import os
import struct
import zlib

import numpy as np
import pytest

from desmo_b200 import shards
from desmo_b200.dist import shard_bounds
from oracle import desmo_oracle as orc


def _vtu(path, vel, fmt):
    """A minimal UnstructuredGrid file with a 3-component point array "velocity" and a decoy scalar, in one of VTK's encodings."""
    npts = vel.shape[0]
    data32 = vel.astype("<f4").tobytes()
    decoy = np.arange(npts, dtype="<f4").tobytes()
    attrs, appended = "", ""

    def arr(name, ncomp, payload_bytes, values):
        nonlocal appended
        if fmt == "ascii":
            return f'<DataArray type="Float32" Name="{name}" NumberOfComponents="{ncomp}" format="ascii">\n' + " ".join(repr(float(v)) for v in values) + "\n</DataArray>"
        if fmt == "binary":
            body = base64.b64encode(struct.pack("<I", len(payload_bytes)) + payload_bytes).decode()
            return f'<DataArray type="Float32" Name="{name}" NumberOfComponents="{ncomp}" format="binary">\n{body}\n</DataArray>'
        if fmt == "binary-sep":  # size header base64-encoded on its own (older writers)
            body = base64.b64encode(struct.pack("<I", len(payload_bytes))).decode() + base64.b64encode(payload_bytes).decode()
            return f'<DataArray type="Float32" Name="{name}" NumberOfComponents="{ncomp}" format="binary">\n{body}\n</DataArray>'
        if fmt == "zlib":
            bs = 1 << 10
            blocks = [payload_bytes[i:i + bs] for i in range(0, len(payload_bytes), bs)]
            comp = [zlib.compress(b) for b in blocks]
            head = struct.pack(f"<{3 + len(blocks)}Q", len(blocks), bs, len(blocks[-1]) % bs, *[len(c) for c in comp])
            body = base64.b64encode(head).decode() + base64.b64encode(b"".join(comp)).decode()
            return f'<DataArray type="Float32" Name="{name}" NumberOfComponents="{ncomp}" format="binary">\n{body}\n</DataArray>'
        off = len(appended.encode("latin-1"))
        appended += (struct.pack("<I", len(payload_bytes)) + payload_bytes).decode("latin-1")
        return f'<DataArray type="Float32" Name="{name}" NumberOfComponents="{ncomp}" format="appended" offset="{off}"/>'

    if fmt == "zlib":
        attrs = ' header_type="UInt64" compressor="vtkZLibDataCompressor"'
    pdata = arr("pressure", 1, decoy, np.arange(npts)) + "\n" + arr("velocity", 3, data32, vel.astype(np.float32).reshape(-1))
    xml = (f'<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian"{attrs}>\n<UnstructuredGrid>\n'
           f'<Piece NumberOfPoints="{npts}" NumberOfCells="0">\n<PointData Vectors="velocity">\n{pdata}\n</PointData>\n</Piece>\n</UnstructuredGrid>\n')
    with open(path, "wb") as fh:
        fh.write(xml.encode("latin-1"))
        if fmt == "appended":
            fh.write(b'<AppendedData encoding="raw">\n_' + appended.encode("latin-1") + b"\n</AppendedData>\n")
        fh.write(b"</VTKFile>\n")


@pytest.mark.parametrize("fmt", ["ascii", "binary", "binary-sep", "zlib", "appended"])
def test_vtu_reader_decodes_every_encoding(tmp_path, fmt):
    rng = np.random.default_rng(1)
    vel = rng.standard_normal((257, 3)).astype(np.float32)
    path = os.path.join(tmp_path, "v.vtu")
    _vtu(path, vel, fmt)
    got = shards.read_vtu_point_array(path, "velocity")
    assert got.shape == (257, 3) and np.array_equal(got.astype(np.float32), vel)
    assert shards.read_vtu_point_array(path, "pressure").shape == (257, 1)
    with pytest.raises(KeyError):
        shards.read_vtu_point_array(path, "f_18")


def test_series_conversion_matches_the_reference_data_matrix(tmp_path):
    """convert_vtu_series + ShardReader reproduce the X of read_velocity_data (CYL:52-68: per-step arrays reshaped to columns, stacked,
    flatten('F'), reshape) and slabs of mesh points partition its rows."""
    rng = np.random.default_rng(2)
    npts, t1, tn = 300, 5, 17
    steps = {i: rng.standard_normal((npts, 3)).astype(np.float32) for i in range(t1, tn)}
    for i, v in steps.items():
        _vtu(os.path.join(tmp_path, f"velocity_{i}.vtu"), v, ["ascii", "binary", "zlib", "appended"][i % 4])
    out = os.path.join(tmp_path, "raw")
    meta = shards.convert_vtu_series(str(tmp_path) + "/", "velocity_", t1, tn, out, steps_per_file=5)
    assert meta["n_points"] == npts and meta["d_in"] == 3 and meta["m_in"] == tn - t1 and len(meta["files"]) == 3
    # the reference's construction, statement for statement (CYL:52-68)
    velocity_list = [np.reshape(steps[i], (-1, 1)) for i in range(t1, tn)]
    X = np.asarray(velocity_list).flatten("F")
    X = np.reshape(X, (-1, tn - t1))
    rd = shards.ShardReader(out)
    assert np.array_equal(rd.data_matrix(), X.astype(np.float64))
    world = 2
    parts = [rd.load_slab(rank, world) for rank in range(world)]
    assert sum(p.shape[1] for p in parts) == npts * 3
    for rank, part in enumerate(parts):
        lo, hi = shard_bounds(npts, world, rank)
        assert np.array_equal(part, X.T[:, lo * 3:hi * 3])  # V = X.T as read, the layout desmo_preprocess consumes


@pytest.mark.gpu
def test_shards_feed_device_preprocess_and_pod(tmp_path):
    """Shards -> slab -> desmo_preprocess -> POD on the device == the reference's numpy pre-processing (CYL:170-187) + SVD (CYL:199)."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(3)
    npts, m_in = 700, 40
    t = np.linspace(0, 6, m_in)
    base = rng.standard_normal((npts, 3, 1)) + 0.5 * rng.standard_normal((npts, 3, 1)) * np.sin(t)[None, None, :] \
        + 0.3 * rng.standard_normal((npts, 3, 1)) * np.cos(2.1 * t)[None, None, :]
    out = os.path.join(tmp_path, "raw")
    shards.write_raw_shards((base[:, :, k].astype(np.float32) for k in range(m_in)), out, steps_per_file=16)
    model, mean, sigma = shards.model_from_shards(out, polyorder=2, r_DESMO=3, d_use=2, device="cuda:0")
    X = shards.ShardReader(out).data_matrix()
    Xp, Xmean = orc.preprocess(X, 3, 2)  # convert3Dto2D_data + convertToMagnitude(X, 2) + subtract_mean
    want = Xp.T.astype(np.float32)
    got = model.engine.U[:, :npts].cpu().numpy()
    ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp == 0).mean() > 0.999
    assert np.allclose(mean.cpu().numpy(), Xmean, rtol=1e-13, atol=1e-15)
    _, _, S, _ = orc.pod_analysis(Xp, 3)
    assert np.linalg.norm(sigma.cpu().numpy() - S[:3]) / np.linalg.norm(S[:3]) < 1e-4
