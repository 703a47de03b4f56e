"""Golden PHOSC labels from the reference's own generator functions (imported from /root/reference; set_phos_version points at
a /cluster path that does not exist here, so its four assignments are repeated with the CSV that lies beside the module):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_phosc.py   ->  tests/golden/phosc_labels.npz"""
import os
import sys

import numpy as np

REF = os.environ.get("WD_REFERENCE_DIR", "/root/reference")
UT = os.path.join(REF, "ResPhoSCNetZSL", "modules", "utils")
sys.path.insert(0, UT)
import phoc_generator as PC  # noqa: E402
import phos_generator as PS  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
WORDS = ["text", "getting", "prop", "a", "it", "The", "Handwriting", "zebra", "QUICK", "jumpsOver", "abcdefghij", "MixedCase",
         "ooo", "w", "xylophone", "Stylist"]


def main():
    csv_path = os.path.join(UT, "Alphabet.csv")
    PS.alphabet_csv = csv_path  # phos_generator.py:39-55 (set_phos_version) with a local path
    PS.alphabet_dict = PS.create_alphabet_dictionary(csv_path)
    PS.csv_num_cols = PS.get_number_of_columns(csv_path)
    npcsv = np.genfromtxt(csv_path, dtype=int, delimiter=",")
    PS.numpy_csv = np.delete(npcsv, 0, 1)
    PC.set_phoc_version("eng")
    labels = []
    for w in WORDS:
        ph = np.asarray(PS.generate_label(w))
        pc = np.asarray(PC.generate_phoc_vector(w), dtype=np.float32)
        labels.append(np.concatenate((ph, pc)))
    lab = np.stack(labels)
    assert lab.shape == (len(WORDS), 769), lab.shape
    np.savez_compressed(os.path.join(OUT, "phosc_labels.npz"), words=np.array(WORDS), labels=lab.astype(np.int32))
    print("phosc labels", lab.shape, "max", lab.max(), "bigram part nonzero:", int(lab[:, -100:].sum()))


if __name__ == "__main__":
    main()
