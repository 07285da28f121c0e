"""The C++ host driver (bin/tidk) end to end: FASTA file in, the reference's output files out,
compared byte for byte with the oracle's text."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIDK = os.path.join(ROOT, "telomeric_identifier_b200", "bin", "tidk")
DB = os.path.join(ROOT, "tests", "golden", "clades_min.csv")


def write_fasta(path, ids, seq, off, width=60, gz=False):
    import gzip
    opener = gzip.open if gz else open
    with opener(path, "wb") as fh:
        for i, rid in enumerate(ids):
            fh.write(f">{rid} synthetic record {i}\n".encode())
            s = seq[int(off[i]):int(off[i + 1])].tobytes()
            for j in range(0, len(s), width):
                fh.write(s[j:j + width] + b"\n")


def run(*args):
    return subprocess.run([TIDK, *args], capture_output=True, text=True, timeout=300)


@pytest.fixture(scope="module")
def genome(tmp_path_factory):
    from telomeric_identifier_b200 import synth
    d = tmp_path_factory.mktemp("cli")
    ids, seq, off = synth.config2(scale=0.01)
    fa = str(d / "g.fa")
    write_fasta(fa, ids, seq, off)
    seqs = [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(ids))]
    return d, fa, ids, seq, off, seqs


@pytest.mark.gpu
def test_cli_search_tsv_and_bedgraph(genome):
    d, fa, ids, seq, off, seqs = genome
    r = run("search", fa, "-s", "ttagg", "-o", "out", "-d", str(d / "s"), "--gpu-stats")
    assert r.returncode == 0, r.stderr
    assert "[+]\tSearching genome for telomeric repeat: ttagg" in r.stderr
    assert f"[+]\tChromosome {ids[0]} processed" in r.stderr and "[+]\tFinished searching genome." in r.stderr
    got = open(d / "s" / "out_telomeric_repeat_windows.tsv").read()
    assert got == O.search_text(ids, seqs, b"ttagg", 10000)
    r = run("search", fa, "--string", "TTAGG", "--window", "3333", "--output", "b", "--dir", str(d / "s"), "-e", "bedgraph")
    assert r.returncode == 0, r.stderr
    assert open(d / "s" / "b_telomeric_repeat_windows.bedgraph").read() == O.search_text(ids, seqs, b"TTAGG", 3333, "bedgraph")


@pytest.mark.gpu
def test_cli_find_clades(genome):
    d, fa, ids, seq, off, seqs = genome
    from telomeric_identifier_b200 import clades
    for clade in ("Lepidoptera", "Hymenoptera", "Coleoptera"):
        r = run("find", fa, "-c", clade, "-o", clade, "-d", str(d / "f"), "--database", DB)
        assert r.returncode == 0, r.stderr
        reps = clades.return_telomere_sequence(clade, DB)
        want = O.find_text(ids, seqs, [x.encode() for x in reps], 10000)
        assert open(d / "f" / f"{clade}_telomeric_repeat_windows.tsv").read() == want
    assert "[+]\tSearching genome for 15 telomeric repeats:" in run("find", fa, "-c", "Hymenoptera", "-o", "h", "-d", str(d / "f"), "--database", DB).stderr


@pytest.mark.gpu
def test_cli_gzip_input(genome, tmp_path):
    d, fa, ids, seq, off, seqs = genome
    gz = str(tmp_path / "g.fa.gz")
    write_fasta(gz, ids[:5], seq, off, width=80, gz=True)
    r = run("search", gz, "-s", "AACCT", "-o", "z", "-d", str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "z_telomeric_repeat_windows.tsv").read() == O.search_text(ids[:5], seqs[:5], b"AACCT", 10000)


@pytest.mark.gpu
def test_cli_explore(genome):
    d, fa, ids, seq, off, seqs = genome
    r = run("explore", fa, "--minimum", "5", "--maximum", "12", "-t", "20", "--distance", "0.05")
    assert r.returncode == 0, r.stderr
    assert r.stdout == O.explore_text(seqs, 0.05, 5, 12, 20, literal=True)
    r = run("explore", fa, "--length", "5", "-t", "100", "--distance", "0.01")
    assert r.returncode == 0 and r.stdout == O.explore_text(seqs, 0.01, 5, 5, 100, literal=True)
    assert run("explore", fa, "-l", "5", "--distance", "0.6").returncode == 1  # explore.rs:41-43


def test_cli_argument_errors(tmp_path):
    """No GPU needed: argument handling fails before the device is touched."""
    import __graft_entry__ as g
    g.build()
    assert run("find", "x.fa", "-c", "NotAClade", "-o", "o", "-d", str(tmp_path), "--database", DB).returncode == 2
    assert run("search", "x.fa", "-o", "o", "-d", str(tmp_path)).returncode == 2
    assert run("explore", "x.fa").returncode == 2  # needs --length or both --minimum and --maximum (main.rs:71-76)
    assert run("plot", "-t", "x.tsv").returncode == 2
    p = run("find", "--print", "--database", DB)
    assert p.returncode == 1 and "Lepidoptera\tAACCT" in p.stderr


@pytest.mark.gpu
def test_cli_log_files_and_plot_schema(genome, tmp_path):
    """--log writes the reference's log text (lib.rs:65-242, date aside), and the TSV deserialises into
    plot's TelomericRepeatRecord (plot.rs:76-96): String, i32, i32, i32, String under the header names."""
    import csv
    import re
    d, fa, ids, seq, off, seqs = genome
    r = run("search", fa, "-s", "TTAGG", "-o", "lg", "-d", str(tmp_path), "--log")
    assert r.returncode == 0 and f"[+]\tLog file written to: {tmp_path}/lg.log" in r.stderr
    text = open(tmp_path / "lg.log").read()
    want = (f"tidk version: 0.2.7\nLog information for output file: {tmp_path}/lg_telomeric_repeat_windows.tsv\n"
            f"Date: DATE\n`tidk search` was run with the following parameters:\n    Input fasta: {fa}\n"
            f"    Telomeric repeat search string: TTAGG\n    Window size: 10000\n                    \n")
    assert re.sub(r"Date: \d{4}-\d\d-\d\d: \d\d:\d\d:\d\d", "Date: DATE", text) == want
    r = run("find", fa, "-c", "Coleoptera", "-o", "fl", "-d", str(tmp_path), "--database", DB, "--log")
    assert r.returncode == 0
    text = open(tmp_path / "fl.log").read()
    assert text.endswith("    Window size: 10000\n    Clade chosen: Coleoptera\n    Telomeric repeats queried: AACCC, AACCT, AACAGACCCG, ACCTG\n")
    assert f"Log information for output file: {tmp_path}/fl_telomeric_repeat_windows.csv\n" in text
    r = subprocess.run([TIDK, "explore", fa, "-m", "5", "-x", "6", "--distance", "0.07", "--log"], capture_output=True,
                       text=True, cwd=str(tmp_path))
    assert r.returncode == 0
    text = open(tmp_path / "tidk-explore.log").read()
    assert text.endswith("    Explored telomeric repeat units of length: 0\n    Or from length: 5\n    To length: 6\n"
                         "    Threshold: 100\n    Searching at 7.000000000000001% distance from chromosome end\n")
    # plot.rs schema
    with open(tmp_path / "lg_telomeric_repeat_windows.tsv", newline="") as fh:
        rows = list(csv.DictReader(fh, delimiter="\t"))
    assert list(rows[0].keys()) == ["id", "window", "forward_repeat_number", "reverse_repeat_number", "telomeric_repeat"]
    for row in rows:
        for col in ("window", "forward_repeat_number", "reverse_repeat_number"):
            assert 0 <= int(row[col]) < 2 ** 31
        assert row["telomeric_repeat"] == "TTAGG" and row["id"] in ids


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 3, 8])
def test_cli_sharded_over_several_contexts(genome, gpus):
    """`--gpus N`: the windows are cut into N contiguous ranges (long records split at multiples of W),
    one context and one thread per range; same files as the oracle's text.  On a one-GPU box every
    range runs on device 0 (TIDK_GPUS_SAME_DEVICE=1), which exercises the plan and the merge."""
    d, fa, ids, seq, off, seqs = genome
    env = dict(os.environ, TIDK_GPUS_SAME_DEVICE="1")
    out = str(d / f"g{gpus}")
    r = subprocess.run([TIDK, "search", fa, "-s", "TTAGG", "-w", "7777", "-o", "s", "-d", out, "--gpus", str(gpus)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    assert open(os.path.join(out, "s_telomeric_repeat_windows.tsv")).read() == O.search_text(ids, seqs, b"TTAGG", 7777)
    from telomeric_identifier_b200 import clades
    reps = clades.return_telomere_sequence("Coleoptera", DB)
    r = subprocess.run([TIDK, "find", fa, "-c", "Coleoptera", "-o", "f", "-d", out, "--database", DB, "--gpus", str(gpus)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    assert open(os.path.join(out, "f_telomeric_repeat_windows.tsv")).read() == O.find_text(ids, seqs, [x.encode() for x in reps], 10000)
    # explore: whole records in N contiguous groups, run lists merged in (k, record) order -- the table is the
    # same text as on one device (the estimator depends on the arrival order, explore.rs:395-423)
    r = subprocess.run([TIDK, "explore", fa, "--minimum", "5", "--maximum", "12", "-t", "20", "--distance", "0.05",
                        "--gpus", str(gpus)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout == O.explore_text(seqs, 0.05, 5, 12, 20, literal=True)


@pytest.mark.gpu
def test_cli_gpus_is_an_upper_bound(genome):
    """`--gpus N` takes up to N devices, one per 4 Gb of work (creating a CUDA context costs more than counting a small
    genome): a small FASTA with `--gpus 8` runs on ONE device -- also on a one-GPU box -- and writes the same file."""
    import json
    d, fa, ids, seq, off, seqs = genome
    out = str(d / "g8_auto")
    env = {k: v for k, v in os.environ.items() if k not in ("TIDK_GPUS_SAME_DEVICE", "TIDK_GPUS_EXACT")}
    r = subprocess.run([TIDK, "search", fa, "-s", "TTAGG", "-w", "7777", "-o", "s", "-d", out, "--gpus", "8", "--gpu-stats"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    stats = json.loads([l for l in r.stderr.splitlines() if l.startswith("{")][-1])
    assert stats["devices_used"] == 1
    assert open(os.path.join(out, "s_telomeric_repeat_windows.tsv")).read() == O.search_text(ids, seqs, b"TTAGG", 7777)
