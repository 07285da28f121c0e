"""Regenerates the fixtures in tests/golden/ from the reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

clades_min.csv: the (Order, Telomeric repeat) columns of clades/curated.csv -- the only two
columns clades.rs:96-141 reads for `tidk find` -- with the other serde columns left empty.
reference_vectors.json: the golden vectors of the reference's inline unit tests,
transcribed with their file:line.
"""
import csv
import json
import os

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

with open(os.path.join(REF, "clades", "curated.csv"), newline="") as fh, \
        open(os.path.join(HERE, "clades_min.csv"), "w", newline="") as out:
    w = csv.writer(out)
    w.writerow(["Phylum", "Order", "Family", "Species", "Telomeric repeat", "Notes", "Ref"])
    for row in csv.DictReader(fh):
        w.writerow(["", row["Order"], "", "", row["Telomeric repeat"], "", ""])

vectors = {
    "search.rs:183-199": {"seq": "TTAGGTTAGGTTAGGCAGCATCACACTGATCATCTGATTAGGTTAGGTTAGG", "query": "TTAGG", "window": 20,
                          "rows": ["test1\t20\t3\t0\tTTAGG", "test1\t40\t0\t0\tTTAGG", "test1\t52\t2\t0\tTTAGG"]},
    "finder.rs:213-235": {"seq": "AAACCCTAAACCCTAAACCCTTGAGAGAGGGGGTGTGGGGAGGGGTTGAGAAACCCT", "repeats": ["AAACCCT"],
                          "window": 20,
                          "rows": ["test1\t20\t2\t0\tAAACCCT", "test1\t40\t0\t0\tAAACCCT", "test1\t57\t1\t0\tAAACCCT"]},
    "utils.rs:178-181": {"in": "AAACCCTTG", "revcomp": "CAAGGGTTT"},
    "utils.rs:204-217": {"rotations": [["TTAGG", "TAGGT", True], ["TTAGG", "AGGTT", True], ["TTAGG", "AACCT", False]]},
    "utils.rs:219-227": {"lex_min": [["TTAGG", "AACCT"], ["TAGGT", "AACCT"]]},
    "utils.rs:231-238": {"motif": "AACCT", "haystack": "AACCTAACCTAACCTAACCTAACCTAACCTAACTAACCT",
                         "indexes": [0, 5, 10, 15, 20, 25, 34]},
    "explore.rs:477-490": {"periods": [["AAAAAAAA", 1], ["ATATATAT", 2], ["AATAATAAT", 3]]},
    "explore.rs:493-626": {"GENOME": "AACCTAACCTAACATATCGTAACCTAACCTAACCTAACCTAACATATCGTAACCTAACCT",
                           "GENOME_2": "AACCTAACCTTAAATTAAATAACCTAACCTAACCTAACCTTAAATTAAATAACCTAACCT",
                           "half": "AACCTAACCTAACATATCGTAACCTAACCT",
                           "chunks": [[0, "AACCT"], [5, "AACCT"], [20, "AACCT"], [25, "AACCT"]],
                           "index_left": [[0, 10, "AACCT"], [20, 30, "AACCT"]],
                           "length_group_5": 3, "estimates": [["AACCT", 4]]},
}
json.dump(vectors, open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=1)
print("wrote", os.listdir(HERE))
