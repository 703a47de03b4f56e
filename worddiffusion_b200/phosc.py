"""PHOSC labels for ``UNetModelPhosc.forward(x, phoscLabels, ...)`` computed on the device (``wd_phosc_tokenize``).

The reference builds this [769] vector per word on the host (ResPhoSCNetZSL/modules/utils/phos_generator.py:59-78 and
phoc_generator.py:17-90; call site trainGWModifyCondition.py:391-405) -- the step immediately in front of the hot path
(SURVEY.md section 8f).  Here the words travel as a small byte tensor and the integer pyramid is built by one kernel."""
import ctypes as C

import torch

from ._lib import WdError, check, lib

PHOSC_LEN = 769


def phosc_labels(words, device="cuda:0", max_len=None):
    """words: list of str -> int32 tensor [len(words), 769] on ``device``.  Spaces and underscores are removed first
    (trainGWModifyCondition.py:394); a character outside a-zA-Z raises KeyError like the reference's alphabet lookup."""
    device = torch.device(device)
    if device.type != "cuda":
        raise WdError("worddiffusion_b200 has no CPU path: phosc_labels needs a CUDA (B200) device")
    clean = [w.replace(" ", "").replace("_", "") for w in words]
    for w in clean:
        for ch in w:
            if not (("a" <= ch <= "z") or ("A" <= ch <= "Z")):
                raise KeyError(ch)
    n = max(1, max((len(w) for w in clean), default=1)) if max_len is None else int(max_len)
    if any(len(w) > n for w in clean):
        raise ValueError(f"a word is longer than max_len = {n}")
    host = torch.zeros((len(clean), n), dtype=torch.uint8)
    for i, w in enumerate(clean):
        if w:
            host[i, : len(w)] = torch.tensor(list(w.encode("ascii")), dtype=torch.uint8)
    dev = host.to(device)
    out = torch.empty((len(clean), PHOSC_LEN), dtype=torch.int32, device=device)
    bad = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().wd_phosc_tokenize(C.c_void_p(dev.data_ptr()), len(clean), n, C.c_void_p(out.data_ptr()),
                                      C.c_void_p(bad.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "wd_phosc_tokenize")
    return out
