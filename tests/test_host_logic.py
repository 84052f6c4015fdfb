"""Host-side logic that needs no GPU: --md parsing, the reader, table packing, result strings,
synthetic tables, sharding, and the N>1 gather under gloo (world_size 2 on CPU)."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import mcz_oracle as orc
from pymcz_b200 import _scales as sc
from pymcz_b200 import mcz, metallicity
from pymcz_b200.synth import synthetic_table, shard_bounds, CONFIGS


def test_keys_and_lines_match_oracle():
    assert sc.KEYS == orc.KEYS and sc.LINES == orc.LINES
    assert metallicity.get_keys() == orc.KEYS and metallicity.get_errkeys() == ['PM14err']
    assert mcz.alllines == orc.LINES


def test_parse_md_tokens():
    m, third, unk = sc.parse_md("all")
    assert m == sc.T_ALL and not third and not unk
    m, third, unk = sc.parse_md("KD02,D13,foo,P05")
    assert m == sc.T_KD02 | sc.T_P05 and third == ["D13"] and unk == ["foo"]
    assert sc.parse_md("KD02comb")[0] == sc.T_KD02          # documented alias (README.md:55-56)
    assert sc.parse_md("M08all")[0] == sc.T_M08ALL            # metallicity.py:243-246
    # `all` omits P05 (metallicity.py:178) and the deprecated scales
    assert not (sc.T_ALL & (sc.T_P05 | sc.T_C01 | sc.T_P01 | sc.T_DP00 | sc.T_D16))


def test_readfile_and_tables(golden_dir):
    """exampledata as shipped by the reference (input/exampledata_{meas,err}.txt), float32 parse."""
    fi = mcz.input_data("exampledata", golden_dir)
    assert fi != -1
    name, flux, err, nm, path, bss = fi
    assert nm == 2 and name == "exampledata"
    assert flux.dtype['Ha'] == np.float32                     # mcz.py:150
    meas_t, err_t = mcz.tables_from_structured(flux, err, nm)
    assert meas_t.shape == (2, 11) and meas_t.dtype == np.float32
    assert np.isclose(meas_t[0, 5], 4.026) and np.isnan(meas_t[0, 4]) and np.isnan(meas_t[1, 2])
    assert err_t[1, 2] == 0.0 and np.isclose(err_t[0, 0], 0.1005)
    assert mcz.input_data("nosuchfile", golden_dir) == -1


def test_savehist_line_matches_oracle_strings(capsys):
    rng = np.random.default_rng(3)
    for _ in range(50):
        data = rng.normal(8.7, 0.05, 300)
        st, med, p16, p84, n = orc.savehist_stats(data)
        s = mcz.savehist_line("x", "KD02comb", sc.ST_OK, med, p16, p84)
        assert s == orc.savehist_string(st, med, p16, p84)
    assert mcz.savehist_line("x", "k", sc.ST_EMPTY, np.nan, np.nan, np.nan) == "-1\t -1\t -1"
    assert mcz.savehist_line("x", "k", sc.ST_NONPOS, np.nan, np.nan, np.nan) == "-1\t -1\t -1"
    mcz.savehist_line("x", "k", sc.ST_OK, 8.69, 8.69, 8.69)
    assert "(no distribution)" in capsys.readouterr().out


def test_synthetic_table_is_reproducible_and_sliceable():
    a_m, a_e = synthetic_table(1000, config=4)
    b_m, b_e = synthetic_table(1000, config=4, start=300, stop=700)
    assert np.array_equal(a_m[300:700], b_m) and np.array_equal(a_e[300:700], b_e)
    assert a_m.dtype == np.float32 and np.all(np.isfinite(a_m)) and np.all(a_m / a_e >= 9.9)
    assert set(CONFIGS) == {3, 4, 5}
    covered = []
    for r in range(8):
        lo, hi = shard_bounds(1_000_003, 8, r)
        covered.append((lo, hi))
    assert covered[0][0] == 0 and covered[-1][1] == 1_000_003
    assert all(covered[i][1] == covered[i + 1][0] for i in range(7))


def test_cli_flags_match_reference(monkeypatch, tmp_path, capsys):
    """Same flags as mcz.py:712-733; nsample gate 1 or >= 10 (mcz.py:762)."""
    os.makedirs(tmp_path / "input")
    mcz.main(["exampledata", "5", "--path", str(tmp_path)])
    assert "nsample must be at least 100" in capsys.readouterr().out
    with pytest.raises(SystemExit):
        mcz.main(["exampledata"])
    with pytest.raises(SystemExit):
        mcz.main(["exampledata", "100", "--binmode", "zz"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, G, q):
    import torch.distributed as td
    from pymcz_b200 import dist
    td.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lo, hi = shard_bounds(G, world, rank)
    # fake per-rank result tables whose content encodes the global galaxy index
    idx = torch.arange(lo, hi, dtype=torch.float32)
    tables = dict(pct=idx[:, None, None].repeat(1, sc.NKEYS, 3), nvalid=idx.to(torch.int32)[:, None].repeat(1, sc.NKEYS),
                  status=(idx.to(torch.int64) % 4).to(torch.int8)[:, None].repeat(1, sc.NKEYS), Z=None)
    out = dist.gather_tables(tables, G)
    assert dist.rank() == rank and dist.world_size() == world
    ok = (out["Z"] is None and out["pct"].shape == (G, sc.NKEYS, 3)
          and torch.equal(out["pct"][:, 0, 0], torch.arange(G, dtype=torch.float32))
          and torch.equal(out["nvalid"][:, 5], torch.arange(G, dtype=torch.int32))
          and torch.equal(out["status"][:, 1], (torch.arange(G) % 4).to(torch.int8)))
    q.put((rank, bool(ok)))
    td.destroy_process_group()


@pytest.mark.parametrize("G", [10, 7, 1])
def test_gather_tables_gloo_world2(G):
    """The N>1 path of the driver: contiguous galaxy shards (also uneven / empty ones) and the single
    all-gather of the result tables, on the gloo backend with two CPU processes."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, G, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
