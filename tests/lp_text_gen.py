"""Seeded generator of random .lp texts inside the dialect the reference reads (test infrastructure).
Used by tests/golden/make_parser_cases.py (committed golden vectors) and by the differential test
against the reference binary when it is present."""
import random


def random_lp_text(seed, m=None, n=None, bounds=True, equalities=True, quirks=True):
    rng = random.Random(seed)
    m = m or rng.randint(2, 9)
    n = n or rng.randint(2, 10)
    names = ["x%d" % (j + 1) for j in range(n)]
    if quirks and rng.random() < 0.3:
        names = [rng.choice(["a", "b", "y", "z", "w", "v"]) + "%d" % j for j in range(n)]
    x0 = [rng.randint(0, 6) for _ in range(n)]          # a feasible point keeps most cases solvable
    lines = []
    if quirks and rng.random() < 0.5:
        lines.append(rng.choice(["\\", "//", "#"]) + " generated case %d" % seed)
    sense = rng.choice(["Maximize", "Minimize", "MAXIMIZE", "min", "max"])
    lines.append(sense)

    def term(coef, name, first):
        if coef == 1 and rng.random() < 0.5:
            txt = name
        elif coef == -1 and rng.random() < 0.5:
            txt = "- " + name
            return txt if first else " " + txt
        else:
            txt = ("%g" % abs(coef)) + rng.choice([" ", ""]) + name
            if coef < 0:
                txt = "- " + txt
        if first:
            return txt
        return (" " if coef < 0 else " + ") + txt

    def fmt_coef():
        c = rng.choice([1, 1, 2, 3, 4, 5, 7, 0.5, 1.5, 2.25, 10, 12])
        return c if rng.random() < 0.75 else -c

    obj = [(fmt_coef(), nm) for nm in names if rng.random() < 0.85] or [(1, names[0])]
    maximize = sense.lower().startswith("max")
    text = ""
    for k, (cf, nm) in enumerate(obj):
        text += term(cf, nm, k == 0)
    lines.append(" " + rng.choice(["obj: ", "cost : ", ""]) + text)
    lines.append(rng.choice(["Subject To", "subject to", "SUBJECT TO"]))
    labels = ["c%d" % (i + 1) for i in range(m)]
    rng.shuffle(labels)
    used = set(nm for _, nm in obj)
    for i in range(m):
        k = rng.randint(1, min(n, 5))
        cols = rng.sample(range(n), k)
        row = [(fmt_coef(), names[j]) for j in cols]
        lhs0 = sum(cf * x0[j] for (cf, _), j in zip(row, cols))
        kind = rng.choice(["<=", "<=", "<=", ">=", "="] if equalities else ["<=", "<=", ">="])
        if kind == "<=":
            rhs = lhs0 + rng.randint(0, 8)
        elif kind == ">=":
            rhs = lhs0 - rng.randint(0, 8)
        else:
            rhs = lhs0
        text = ""
        for t, (cf, nm) in enumerate(row):
            text += term(cf, nm, t == 0)
            used.add(nm)
        rel = kind if not quirks else rng.choice([kind, kind, kind.replace("<=", "<").replace(">=", ">")])
        if quirks and len(text) > 12 and rng.random() < 0.2:
            cut = text.rfind(" ", 0, len(text) // 2 + 4)
            if cut > 0:
                lines.append(" %s: %s" % (labels[i], text[:cut]))
                lines.append("      %s %s %g" % (text[cut:], rel, rhs))
                continue
        lines.append(" %s: %s %s %g" % (labels[i], text, rel, rhs))
    # keep the LP bounded: an objective-improving direction is capped by a bound or a row
    bound_lines = []
    if bounds:
        for j, nm in enumerate(names):
            if nm not in used:
                continue
            r = rng.random()
            if r < 0.35:
                bound_lines.append(" %s <= %g" % (nm, x0[j] + rng.randint(1, 6)))
            elif r < 0.55:
                lo = max(0, x0[j] - rng.randint(0, 3))
                bound_lines.append(" %g <= %s <= %g" % (lo, nm, x0[j] + rng.randint(1, 6)))
            elif r < 0.65:
                bound_lines.append(" %g <= %s" % (max(0, x0[j] - rng.randint(0, 2)), nm))
    if bound_lines:
        lines.append(rng.choice(["Bounds", "bounds", "BOUNDS"]))
        lines.extend(bound_lines)
    lines.append(rng.choice(["End", "end", "END"]))
    return "\n".join(lines) + "\n"
