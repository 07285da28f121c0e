"""The oracle against EVERY golden vector in the reference's own unit tests
(SURVEY.md section 4).  Each test names the reference test it reproduces."""
from oracle import oracle as O
from oracle import pymodel as M

GENOME = "AACCTAACCTAACATATCGTAACCTAACCTAACCTAACCTAACATATCGTAACCTAACCT"    # explore.rs:493
GENOME_2 = "AACCTAACCTTAAATTAAATAACCTAACCTAACCTAACCTTAAATTAAATAACCTAACCT"  # explore.rs:495
HALF = b"AACCTAACCTAACATATCGTAACCTAACCT"


def test_search_test_search_1():  # search.rs:183-199
    seq = b"TTAGGTTAGGTTAGGCAGCATCACACTGATCATCTGATTAGGTTAGGTTAGG"
    text = O.search_text(["test1"], [seq], b"TTAGG", 20)
    rows = text.splitlines()[1:]
    assert rows == ["test1\t20\t3\t0\tTTAGG", "test1\t40\t0\t0\tTTAGG", "test1\t52\t2\t0\tTTAGG"]
    assert M.search_rows(seq.decode(), "test1", "TTAGG", 20) == rows


def test_finder_test_search_1():  # finder.rs:213-235
    seq = b"AAACCCTAAACCCTAAACCCTTGAGAGAGGGGGTGTGGGGAGGGGTTGAGAAACCCT"
    text = O.find_text(["test1"], [seq], [b"AAACCCT"], 20)
    rows = text.splitlines()[1:]
    assert rows == ["test1\t20\t2\t0\tAAACCCT", "test1\t40\t0\t0\tAAACCCT", "test1\t57\t1\t0\tAAACCCT"]
    assert M.find_rows(seq.decode(), "test1", ["AAACCCT"], 20) == rows


def test_utils_revcomp1():  # utils.rs:178-181
    assert O.revcomp(b"AAACCCTTG") == b"CAAGGGTTT"
    assert M.reverse_complement("AAACCCTTG") == "CAAGGGTTT"


def test_utils_string_rotation():  # utils.rs:204-217
    assert O.string_rotation(b"TTAGG", b"TAGGT")
    assert O.string_rotation(b"TTAGG", b"AGGTT")
    assert not O.string_rotation(b"TTAGG", b"AACCT")


def test_utils_lex_min():  # utils.rs:219-227
    assert O.lex_min(b"TTAGG") == b"AACCT"
    assert O.lex_min(b"TAGGT") == b"AACCT"
    assert M.lex_min("TTAGG") == "AACCT" and M.lex_min("TAGGT") == "AACCT"


def test_utils_motifs1():  # utils.rs:231-238
    hay = b"AACCTAACCTAACCTAACCTAACCTAACCTAACTAACCT"
    assert O.find_motifs(b"AACCT", hay) == [0, 5, 10, 15, 20, 25, 34]


def test_explore_check_repeat_period():  # explore.rs:477-490
    assert O.period(b"AAAAAAAA") == 1 and M.check_repeats("AAAAAAAA") == 1
    assert O.period(b"ATATATAT") == 2 and M.check_repeats("ATATATAT") == 2
    assert O.period(b"AATAATAAT") == 3 and M.check_repeats("AATAATAAT") == 3


def test_explore_split_left_right():  # explore.rs:509-520
    d = O.split_dist(len(GENOME), 0.5)
    g = GENOME.encode()
    assert g[:d] == HALF and g[len(g) - d:] == HALF
    assert M.split_seq_by_distance(GENOME, 0.5) == (HALF.decode(), HALF.decode())


def test_explore_chunks_left_right():  # explore.rs:528-584
    want = [(0, b"AACCT"), (5, b"AACCT"), (20, b"AACCT"), (25, b"AACCT")]
    assert O.chunk_fasta(HALF, 5) == want
    assert M.chunk_fasta(HALF.decode(), 5) == [(p, s.decode()) for p, s in want]


def test_explore_index_left():  # explore.rs:592-612
    runs = O.slice_runs(HALF, 5, 0)
    assert [(s, e, l) for s, e, l, _ in runs] == [(0, 10, b"AACCT"), (20, 30, b"AACCT")]
    assert [f for *_, f in runs] == [1, 0]


def test_explore_get_length_groups():  # explore.rs:614-619
    left = GENOME_2.encode()[:30]
    runs = O.slice_runs(left, 5, 0)
    # AACCT 0-10, TAAAT 10-20, AACCT 20-30
    assert [(s, e, l) for s, e, l, _ in runs] == [(0, 10, b"AACCT"), (10, 20, b"TAAAT"), (20, 30, b"AACCT")]


def test_explore_get_telomeric_repeat_estimates():  # explore.rs:622-626
    left = GENOME_2.encode()[:30]
    runs = O.slice_runs(left, 5, 0)
    labels = [l for _, _, l, _ in runs]
    counts = [(e - s) // 5 for s, e, _, _ in runs]
    assert O.estimates(labels, counts, literal=True) == [(b"AACCT", 4)]
    assert O.estimates(labels, counts, literal=False) == [(b"AACCT", 4)]
    m = M.calculate_indexes(M.chunk_fasta(left.decode(), 5), 5, 0)
    assert M.get_telomeric_repeat_estimates(m) == [("AACCT", 4)]
