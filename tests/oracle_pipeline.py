"""A tiny interpreter of genodsp command lines on top of the plain-C oracle (CPU).  Test
infrastructure only: it lets the CPU test-suite check the oracle against the golden fixtures
(tests/golden/, produced by the reference binary) on machines without /root/reference."""
import math
import os

import numpy as np

from checkers import Oracle

DBL_MAX = float(np.finfo(np.float64).max)


def _num(s):
    return {"inf": DBL_MAX, "+inf": DBL_MAX, "-inf": -DBL_MAX}.get(s, None) if s in ("inf", "+inf", "-inf") else float(s)


def _unit(s):
    mult = {"K": 1000, "M": 1000000, "G": 1000000000}.get(s[-1].upper(), 1)
    return int(float(s[:-1]) * mult + 0.5) if mult != 1 else int(s)


def fmt(v, precision):
    return "%.*f" % (precision, v)


class Pipeline:
    def __init__(self, cwd):
        self.cwd = cwd
        self.orc = Oracle()
        self.vars = {}
        self.stderr = []

    def read_chroms(self, path):
        self.chroms = []
        for line in open(os.path.join(self.cwd, path)):
            f = line.split()
            if f and not f[0].startswith("#"):
                self.chroms.append((f[0], int(f[1])))
        self.sorted = sorted(range(len(self.chroms)), key=lambda i: -self.chroms[i][1])
        self.v = {n: np.zeros(l) for n, l in self.chroms}

    def read_intervals(self, path, val_col, origin=0):
        by = {n: ([], [], []) for n, _ in self.chroms}
        for line in open(os.path.join(self.cwd, path)):
            if line.startswith("track ") or not line.strip() or line.lstrip().startswith("#"):
                continue
            f = line.split()
            if f[0] in by:
                by[f[0]][0].append(int(f[1]) - origin); by[f[0]][1].append(int(f[2]))
                by[f[0]][2].append(1.0 if val_col < 0 else float(f[val_col]))
        return by

    def per_chrom(self, fn):
        for i in self.sorted:
            n = self.chroms[i][0]
            fn(self.v[n])

    def run(self, argv, stdin):
        opts = {"val_col": 3, "precision": 0, "collapse": True, "show": 0, "origin": 0, "window": None}
        ops, cur = [], None
        for a in argv:
            if a == "=":
                cur = []; ops.append(cur)
            elif cur is not None:
                cur.append(a)
            elif a.startswith("--chromosomes="):
                self.read_chroms(a.split("=", 1)[1])
            elif a == "--novalue":
                opts["val_col"] = -1
            elif a.startswith("--precision="):
                opts["precision"] = int(a.split("=")[1])
            elif a == "--nocollapse":
                opts["collapse"] = False
            elif a == "--uncovered:show":
                opts["show"] = 1
            elif a == "--uncovered:NA":
                opts["show"] = -1
            elif a == "--origin=one":
                opts["origin"] = 1
            else:
                raise ValueError(a)
        iv = self.read_intervals(stdin, opts["val_col"], opts["origin"])
        for n, _ in self.chroms:
            s, e, val = iv[n]
            self.orc.accumulate(self.v[n], s, e, None if opts["val_col"] < 0 else val)
        for op in ops:
            self.apply(op, opts)
        return self.report(opts)

    def kw(self, args, names, default=None):
        for a in args:
            for nm in names:
                if a.startswith(nm):
                    return a[len(nm):]
        return default

    def apply(self, op, opts):
        name, args = op[0], op[1:]
        o = self.orc
        # an operator's interval file is read with the global value column unless it says otherwise
        # (so under --novalue every file value is 1)
        fcol = opts["val_col"]
        if "--novalue" in args:
            fcol = -1
        elif self.kw(args, ["--value="]) is not None:
            fcol = int(self.kw(args, ["--value="])) - 1
        pos = [a for a in args if not a.startswith("--")]
        if name == "sum":
            W = _unit(self.kw(args, ["--window="], "100"))
            d = self.kw(args, ["--denom="], "1")
            self.per_chrom(lambda v: o.block_sum(v, W, float(W) if d in ("W", "window") else float(d), d == "actual", 0.0))
        elif name == "slidingsum":
            W = _unit(self.kw(args, ["--window="], "100"))
            d = self.kw(args, ["--denom="], "1")
            self.per_chrom(lambda v: o.sliding_sum(v, W, float(W) if d in ("W", "window") else float(d)))
        elif name == "smooth":
            W = _unit(self.kw(args, ["--window="], "101")); W += (W % 2 == 0)
            self.per_chrom(lambda v: o.smooth(v, W))
        elif name == "cumulativesum":
            self.per_chrom(o.cumulative)
        elif name in ("localmax", "localmin"):
            N = _unit(self.kw(args, ["--neighborhood="], "3"))
            fill = float(self.kw(args, ["--zero=", "--infinity="], "0" if name == "localmax" else repr(DBL_MAX)))
            self.per_chrom(lambda v: o.local_extrema(v, N, name == "localmax", fill))
        elif name in ("bestmax", "bestmin"):
            W = _unit(self.kw(args, ["--window="], "100"))
            self.per_chrom(lambda v: o.best_extrema(v, W, name == "bestmax"))
        elif name == "binarize":
            var = self.kw(args, ["--threshold="])
            if var is not None:
                T = self.vars[var]
                self.stderr.append("[binarize] using %s = %f as threshold" % (var, T))
            else:
                T = float(pos[0]) if pos else 0.0
            self.per_chrom(lambda v: o.binarize(v, T))
        elif name == "addconst":
            c = float(args[0])
            self.per_chrom(lambda v: o.addconst(v, c))
        elif name == "abs":
            self.per_chrom(o.abs)
        elif name == "clip":
            mn, mx = self.kw(args, ["--min="]), self.kw(args, ["--max="])
            self.per_chrom(lambda v: o.clip(v, None if mn is None else float(mn), None if mx is None else float(mx)))
        elif name == "erase":
            mn, mx = self.kw(args, ["--min="]), self.kw(args, ["--max="])
            self.per_chrom(lambda v: o.erase(v, None if mn is None else float(mn), None if mx is None else float(mx),
                                             "--keep:inside" in args, 0.0))
        elif name == "invert":
            if pos:
                mid = float(pos[0])
            else:
                allv = np.concatenate([self.v[n] for n, _ in self.chroms])
                mid = (allv.min() + allv.max()) / 2.0
            self.per_chrom(lambda v: o.invert(v, mid))
        elif name in ("open", "close", "dilate", "erode"):
            L = _unit(pos[0]); T = float(self.kw(args, ["--threshold="], "0"))
            if name == "open":
                self.per_chrom(lambda v: o.open(v, L, T))
            elif name == "close":
                self.per_chrom(lambda v: o.close(v, L, T))
            elif name == "dilate":
                self.per_chrom(lambda v: o.dilate(v, L // 2, L - L // 2, T))
            else:
                self.per_chrom(lambda v: o.erode(v, L // 2, L - L // 2, T))
        elif name in ("clump", "anticlump"):
            T = float(pos[0]); L = _unit(self.kw(args, ["--length="], "100"))
            one = float(self.kw(args, ["--one="], "1"))
            self.per_chrom(lambda v: o.clump(v, T, L, name == "clump", one, 0.0))
        elif name == "percentile":
            # values by sorting the qualifying samples; post-state by tests/percentile_model.py (pinned to the
            # reference by tests/test_oracle_percentile_state.py)
            import percentile_model as pmodel
            spec_p = pos[0]
            prec = int(self.kw(args, ["--precision="], str(opts["precision"])))
            stride = int(_unit(self.kw(args, ["--window=", "W="], "1")))
            mn = _num(self.kw(args, ["--min="], "-inf")); mx = _num(self.kw(args, ["--max="], "inf"))
            quiet = "--quiet" in args or "--silent" in args
            if ".." in spec_p:
                lo_s, rest = spec_p.split("..")
                hi_s, step_s = (rest.split("by") + ["1"])[:2]
                plist = []
                lo_m, hi_m, st_m = int(round(float(lo_s) * 1000)), int(round(float(hi_s) * 1000)), int(round(float(step_s) * 1000))
                pmv = lo_m
                while pmv <= hi_m:
                    plist.append(pmv); pmv += st_m
            else:
                plist = [int(round(float(spec_p) * 1000))]
            names = [self.chroms[i][0] for i in self.sorted]
            lengths = [self.v[n].size for n in names]
            cat = np.concatenate([self.v[n] for n in names])
            q = pmodel.qualifying_mask(lengths, cat, stride, mn, mx)
            srt = np.sort(cat[q])
            rank = 0
            for pmv in plist:
                rank = srt.size - 1 if pmv >= 100000 else o.percentile_rank(srt.size, pmv)
                val = srt[rank]
                label = ("%d" % (pmv // 1000)) if pmv % 1000 == 0 else ("%g" % (pmv / 1000.0))
                self.vars["percentile" + label] = val
                if not quiet:
                    self.stderr.append("percentile %.3f is %s" % (pmv / 1000.0, fmt(val, prec)))
            state, _n = pmodel.post_state(lengths, cat, rank, stride, mn, mx)
            at = 0
            for n in names:
                self.v[n] = state[at:at + self.v[n].size].copy(); at += self.v[n].size
        elif name == "input":
            # op_input_apply (opio.c:225-256) -> read_intervals with clear (genodsp.c:1307-1330): every vector
            # starts at missingVal (an int, opio.c:35) and takes the intervals one after the other, cell by cell
            icol = opts["val_col"]
            if "--novalue" in args or "--novalues" in args or "--value=none" in args:
                icol = -1
            elif self.kw(args, ["--value="]) is not None:
                icol = int(self.kw(args, ["--value="])) - 1
            origin = opts["origin"]
            if "--origin=one" in args or "--origin=1" in args:
                origin = 1
            if "--origin=zero" in args or "--origin=0" in args:
                origin = 0
            missing = float(int(float(self.kw(args, ["--missing="], "0"))))
            ov = self.kw(args, ["--overlap="], "sum")[:3]
            iv = self.read_intervals(pos[0], icol, origin)
            for n, _ in self.chroms:
                v = self.v[n]
                v[:] = missing
                for s, e, val in zip(*iv[n]):
                    s, e = max(s, 0), min(e, v.size)
                    if s >= e:
                        continue
                    seg = v[s:e]
                    fresh = seg == missing
                    with np.errstate(invalid="ignore"):
                        if ov == "min":
                            seg[:] = np.where(fresh | (val < seg), val, seg)
                        elif ov == "max":
                            seg[:] = np.where(fresh | (val > seg), val, seg)
                        else:
                            seg[:] = np.where(fresh, val, seg + val)
        elif name in ("add", "subtract", "multiply", "divide", "and", "masknot"):
            iv = self.read_intervals(pos[0], -1 if name == "masknot" else fcol)    # masknot reads no value column (mask.c:533)
            for n, _ in self.chroms:
                s, e, val = iv[n]
                keep = [i for i, x in enumerate(val) if x != 0.0]                  # val==0 lines are skipped
                s = [s[i] for i in keep]; e = [e[i] for i in keep]; val = [val[i] for i in keep]
                if name in ("add", "subtract"):
                    o.add_intervals(self.v[n], s, e, val, 1.0 if name == "add" else -1.0)
                elif name == "multiply":
                    o.sorted_intervals(self.v[n], s, e, val, 0, 0.0)
                elif name == "divide":
                    o.sorted_intervals(self.v[n], s, e, val, 1, _num(self.kw(args, ["--infinity="], "inf")))
                elif name == "masknot":
                    o.sorted_intervals(self.v[n], s, e, val, 2, float(self.kw(args, ["--mask=", "M="], "0")))
                else:
                    o.sorted_intervals(o.logical_prep(self.v[n]), s, e, val, 3, 0.0)
        elif name in ("mask", "or"):
            iv = self.read_intervals(pos[0], -1 if name == "mask" else fcol)
            for n, _ in self.chroms:
                s, e, val = iv[n]
                if name == "mask":
                    o.mask_intervals(self.v[n], s, e, float(self.kw(args, ["--mask=", "M="], "0")))
                else:
                    o.or_intervals(o.logical_prep(self.v[n]), s, e, val)
        elif name in ("minover", "maxover"):
            iv = self.read_intervals(pos[0], -1)
            fill = _num(self.kw(args, ["--infinity="], "inf")) if name == "minover" else float(self.kw(args, ["--zero="], "0"))
            for n, _ in self.chroms:
                s, e, _val = iv[n]
                self.v[n] = o.over_intervals(self.v[n], s, e, name == "maxover", fill)
        elif name in ("minwith", "maxwith"):
            iv = self.read_intervals(pos[0], fcol)
            for n, _ in self.chroms:
                s, e, val = iv[n]
                self.v[n] = o.with_intervals(self.v[n], s, e, val, name == "maxwith")
        elif name == "map":
            pts = []
            for line in open(os.path.join(self.cwd, pos[0])):
                f = line.split()
                if f and not f[0].startswith("#"):
                    pts.append((float(f[0]), float(f[1])))
            pts.sort()
            vin = np.array([a for a, _ in pts]); vout = np.array([b for _, b in pts])
            for n, _ in self.chroms:
                self.v[n] = o.map_values(self.v[n], vin, vout)
        else:
            raise ValueError("oracle pipeline: operator %s not supported" % name)

    def report(self, opts):
        out = []
        o = opts["origin"]
        for n, l in self.chroms:
            rs, re, rv = self.orc.runs(np.ascontiguousarray(self.v[n]), opts["collapse"], opts["show"])
            prev = 0
            for s, e, x in zip(rs, re, rv):
                if opts["show"] == -1 and s != prev:
                    out.append("%s\t%d\t%d\tNA" % (n, prev + o, s))
                out.append("%s\t%d\t%d\t%s" % (n, s + o, e, fmt(x, opts["precision"])))
                prev = int(e)
            if opts["show"] == -1 and prev != l:
                out.append("%s\t%d\t%d\tNA" % (n, prev + o, l))
        return "\n".join(out) + ("\n" if out else "")
