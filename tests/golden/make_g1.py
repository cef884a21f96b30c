"""Extracts golden vector G1 from the reference notebook's STORED OUTPUTS.

Run in the build container (needs /root/reference):  python tests/golden/make_g1.py
Source: /root/reference/creating images.ipynb, cell 3 (dfm.head(): LOBSTER AMZN 2012-06-21
messages 0-4) and cell 4 (dfo.head(): L2 depth-10 order book AFTER each message).
Also records worked example G2 inputs (gymnax_exchange/jaxob/jorderbook.py:296-302).
The output JSON is committed; nothing on the GPU box reads /root/reference.
"""
import json
import os
import re

NB = "/root/reference/creating images.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "g1_lobster_amzn.json")


def main():
    nb = json.load(open(NB))
    # cell 4: full 40-column rows live in the datawrangler JSON output
    book_rows = None
    for o in nb["cells"][4]["outputs"]:
        v = o.get("data", {}).get("application/vnd.microsoft.datawrangler.viewer.v0+json")
        if v:
            cols = [c["name"] for c in v["columns"]][1:]
            book_rows = [[int(x) for x in r[1:]] for r in v["rows"]]
    assert book_rows and len(book_rows) == 5 and all(len(r) == 40 for r in book_rows)
    # cell 3: messages (time, event_type, order_id, size, price, direction)
    msgs = None
    for o in nb["cells"][3]["outputs"]:
        v = o.get("data", {}).get("application/vnd.microsoft.datawrangler.viewer.v0+json")
        if v:
            msgs = [r[1:] for r in v["rows"]]
    if msgs is None:
        txt = "".join(nb["cells"][3]["outputs"][0]["data"]["text/plain"])
        msgs = [re.split(r"\s+", ln.strip())[1:] for ln in txt.splitlines()[1:]]
    assert len(msgs) == 5
    out = {
        "source": "creating images.ipynb cells 3,4 (stored outputs); LOBSTER AMZN 2012-06-21 level 10",
        "book_columns": cols,
        "messages_lobster": [[str(x) for x in m] for m in msgs],   # time kept as string (exact)
        "book_rows": book_rows,
        "g2": {
            "source": "gymnax_exchange/jaxob/jorderbook.py:296-302",
            "l2init": [354200, 452, 350100, 89, 361200, 100, 344000, 400, 362900, 100, 343100, 100, 364000, 400,
                       338700, 100, 371900, 1100, 337100, 1000, 372200, 100, 336400, 1000, 372300, 200, 336000, 300,
                       372800, 1000, 333600, 1000, 374600, 1000, 332500, 100, 376700, 100, 331600, 100],
            "msgs": [[1, 1, 99, 346000, 8888, 8888, 3400, 5000000], [1, -1, 2, 346000, 8777, 8777, 3401, 5060000]],
        },
    }
    json.dump(out, open(OUT, "w"))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
