#!/usr/bin/env python
"""Drop-in proof, step 1: produce the generated registry file `lib/nnc/cmd/ccv_nnc_cmd.inc` that
`build-cmd.rb` would write for a tree that carries the SM100 backend.  Ruby is not in this image, so the file is derived
from the reference's committed .inc (read where it lies, never copied into this repository):

  * `backend_init_map` and `_ccv_nnc_cmd_backend_ph` are replaced by the 8-backend ones (integration/ccv_nnc_cmd_backend.inc,
    found by the generator's own search restated in tools/gen_backend_ph.py);
  * every existing `_register_command_X_backend_Y(&(init_map[i].backends[j]))` call keeps its command slot i and gets the slot j
    the NEW backend hash assigns to Y (all seven move: the 7-slot hash `(backend >> 15) % 7` no longer applies);
  * the SM100 prototypes and calls (integration/ccv_nnc_cmd_sm100_init.inc) are appended under HAVE_CUDA;
  * the reference's own CUDA registrations stay under HAVE_CUDA but can be configured out with CCV_NNC_DROPIN_SM100_ONLY -- the
    equivalent of a `CUDA_CMD_SRCS` (lib/nnc/cmd/config.mk) that lists only the SM100 sources, which makes SM100 the ONLY
    GPU backend: every GPU command the reference's upper layers issue must then be served by this backend.

  python integration/patch_cmd_inc.py /root/reference/lib/nnc/cmd/ccv_nnc_cmd.inc  out/ccv_nnc_cmd.inc
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(src, dst):
    text = open(src).read()
    backend_inc = open(os.path.join(HERE, "ccv_nnc_cmd_backend.inc")).read()
    sm100_inc = open(os.path.join(HERE, "ccv_nnc_cmd_sm100_init.inc")).read()
    order = re.findall(r'\{\.name = "(CCV_NNC_BACKEND_[A-Z0-9_]+)", \.backend = 0x[0-9a-f]+\}', backend_inc)
    assert len(order) == 8, order
    slot = {name: i for i, name in enumerate(order)}
    # 1. the backend side
    a = text.index("static ccv_nnc_cmd_backend_init_t backend_init_map[] = {")
    b = text.index("};", a) + 3
    init_map8 = backend_inc[backend_inc.index("static ccv_nnc_cmd_backend_init_t"):backend_inc.index("};") + 3]
    text = text[:a] + init_map8 + text[b:]
    m = re.search(r"static inline int _ccv_nnc_cmd_backend_ph\(const uint32_t backend\)\s*\{.*?\n\}\n", text, re.S)
    ph8 = backend_inc[backend_inc.index("static inline int _ccv_nnc_cmd_backend_ph"):]
    text = text[:m.start()] + ph8 + text[m.end():]
    # 2. re-slot the existing registration calls
    n = [0]

    def reslot(mm):
        n[0] += 1
        return "%s(&(init_map[%s].backends[%d]));" % (mm.group(1), mm.group(3), slot[mm.group(2)])
    text = re.sub(r"(_register_command_CCV_NNC_[A-Z0-9_]+_backend_(CCV_NNC_BACKEND_[A-Z0-9_]+))\(&\(init_map\[(\d+)\]\.backends\[\d+\]\)\);", reslot, text)
    assert n[0] > 300, n
    # 3. the reference's CUDA backends become optional; SM100 is added
    text = text.replace("#ifdef HAVE_CUDA", "#if defined(HAVE_CUDA) && !defined(CCV_NNC_DROPIN_SM100_ONLY)")
    protos, calls = sm100_inc.split("static inline void _ccv_nnc_cmd_init_sm100(void)")
    # prototypes go in front of _ccv_nnc_cmd_init(); the helper with the calls too, invoked at the end of _ccv_nnc_cmd_init()
    k = text.index("static inline void _ccv_nnc_cmd_init(void)")
    text = text[:k] + "#ifdef HAVE_CUDA\n" + protos + "static inline void _ccv_nnc_cmd_init_sm100(void)" + calls + "#endif\n\n" + text[k:]
    end = text.rindex("}")
    text = text[:end] + "#ifdef HAVE_CUDA\n\t_ccv_nnc_cmd_init_sm100();\n#endif\n" + text[end:]
    os.makedirs(os.path.dirname(os.path.abspath(dst)), exist_ok=True)
    if os.path.islink(dst):
        os.unlink(dst)
    open(dst, "w").write(text)
    print("patched %d registration calls -> %s" % (n[0], dst))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
