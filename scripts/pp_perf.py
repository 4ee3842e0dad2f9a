import sys, json
for line in sys.stdin:
    try:
        name, js = line.split(" ", 1); j = json.loads(js)
    except Exception:
        print(line.rstrip()[:300]); continue
    print(name, j["shape"], j["bound"], "roof", round(j["roofline_ms"],4), "best", round(j["best_ms"],4), "frac", round(j["frac_of_roofline"],3), "qps", int(j["qps"]), "S", j["stats"]["slices"], "items", j["stats"]["items"])
    for k, v in j["runs"].items():
        for x in v:
            c = x["chain"]
            print("   ", k, "total", round(x["ms"]*1e3,1), "us |", " ".join(f"{n[2:]}:{c[n]['ms']*1e3:.1f}(+{c[n]['gap_ms']*1e3:.1f})" for n in ("k_prep_rows","k_score_topk","k_select_rerank","k_exact_fallback")))
