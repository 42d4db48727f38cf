"""TEST INFRASTRUCTURE: an Engine look-alike backed by the CPU oracle, so that the host-side driver (`bdf_b200.macau`) — loop order,
posterior averaging, RMSE / dump bookkeeping, alpha and multi-relation plumbing — can be exercised by the `-m "not gpu"` suite.
It is not part of the product and is never imported by it (the product has no CPU path); it lives under tests/ for that reason."""
import numpy as np

from oracle import oracle as orc


class OracleEngine:
    def __init__(self, num_latent, seed=0):
        self.D = num_latent
        self.rng = np.random.default_rng(seed)
        self.U, self.rels, self.alpha, self.mean, self.stats = [], [], [], [], {}
        self.launches = 0
        self.calls = []

    def set_seed(self, seed):
        self.rng = np.random.default_rng(seed)

    def add_entity(self, count):
        self.U.append(np.zeros((count, self.D)))
        return len(self.U) - 1

    def add_relation(self, entities, ids, vals):
        self.rels.append((list(entities), np.asarray(ids, dtype=np.int64), np.asarray(vals, dtype=np.float64)))
        self.alpha.append(1.0)
        self.mean.append(0.0)
        return len(self.rels) - 1

    def set_relation_params(self, rel, alpha, mean):
        self.alpha[rel], self.mean[rel] = float(alpha), float(mean)

    def set_factors(self, e, U):
        self.U[e] = np.array(U, dtype=np.float64)

    def get_factors(self, e):
        return self.U[e].copy()

    def sample_mode(self, e, mu, Lambda, z):
        """sample_user2_all!: every row of entity e from all relations it takes part in (src/sampling.jl:251-289)."""
        self.calls.append(("sample", e))
        N = self.U[e].shape[0]
        Z = self.rng.standard_normal((N, self.D)) if z is None else z
        out = np.empty_like(self.U[e])
        uses = [(r, ents.index(e)) for r, (ents, _, _) in enumerate(self.rels) if e in ents]
        if len(uses) == 1:
            # one relation: the multi-worker half-sweep of the C oracle (sample_latent_all2!, src/sampling.jl:149-172) — same
            # arithmetic as the per-row call below, OpenMP over cyclic shards, needed for MovieLens-sized tables
            r, m = uses[0]
            ents, ids, vals = self.rels[r]
            self._idf = getattr(self, "_idf", {})
            if r not in self._idf:
                self._idf[r] = orc.FastIDF(ids, vals, [self.U[k].shape[0] for k in ents])
            Us = [self.U[k] for k in ents]
            Us[m] = out
            orc.sample_latent_all(self._idf[r], m, Us, self.alpha[r], self.mean[r], mu, Lambda, np.ascontiguousarray(Z),
                                  nshards=max(1, min(orc.max_threads(), 16)))
            self.U[e] = out
            return
        order = {r: np.argsort(self.rels[r][1][:, m], kind="stable") for r, m in uses}
        for i in range(N):
            rl = []
            for r, m in uses:
                ents, ids, vals = self.rels[r]
                o = order[r]
                lo, hi = np.searchsorted(ids[o, m], [i + 1, i + 2])
                sel = o[lo:hi]
                others = [k for k in range(len(ents)) if k != m]
                rl.append({"U": [self.U[ents[k]] for k in others], "ids": [ids[sel, k] for k in others], "vals": vals[sel],
                           "offset": self.mean[r], "alpha": self.alpha[r]})
            out[i] = orc.sample_row(self.D, rl, mu if np.ndim(mu) == 1 else mu[i], Lambda, Z[i])
        self.U[e] = out

    def nw_stats(self, e):
        self.calls.append(("stats", e))
        self.stats[e] = orc.nw_stats(self.U[e])
        return self.stats[e]

    def nw_sample(self, e, mu0, b0, Tinv, nu, bartlettA=None, z=None):
        self.calls.append(("draw", e))
        N, NU, NS = self.stats[e]
        mu_N, beta_N, T_N, nu_N = orc.cond_normal_wishart(N, NU, NS, np.asarray(mu0), b0, np.asarray(Tinv), nu)
        A = orc.bartlett_factor(self.rng, self.D, nu_N) if bartlettA is None else bartlettA
        zz = self.rng.standard_normal(self.D) if z is None else z
        return orc.nw_rand(mu_N, beta_N, T_N, A, zz)

    # ---- the deferred / device-resident forms the product loop uses (here they simply run at once on the host) ---------------------
    def set_async(self, on=True):
        pass

    def step_nw_stats(self, e):
        self.nw_stats(e)

    def nw_sample_async(self, e, mu0, b0, Tinv, nu, bartlettA=None, z=None):
        self._draws = getattr(self, "_draws", {})
        self._draws[e] = self.nw_sample(e, mu0, b0, Tinv, nu, bartlettA, z)

    def nw_sample_fetch(self, e):
        return self._draws.pop(e)

    # ---- test set and posterior accumulators (bdf_set_test / bdf_predict_accumulate / bdf_get_test_predictions): src/macau.jl:143-200 ----
    def set_test(self, rel, ids, vals, test_F=None, class_cut=0.0):
        self._test = getattr(self, "_test", {})
        n = len(vals)
        self._test[rel] = {"ids": np.asarray(ids), "vals": np.asarray(vals, dtype=np.float64), "F": test_F, "cut": class_cut,
                           "all": np.zeros(n), "sq": np.zeros(n), "counter": 0}

    def predict_accumulate(self, rel, posterior, clamp=()):
        t = self._test[rel]
        probe_rat = self.predict(rel, t["ids"], t["F"])
        if not posterior:
            t["all"] = probe_rat
        elif t["counter"] == 0:
            t["all"], t["sq"], t["counter"] = probe_rat.copy(), probe_rat ** 2, 1
        else:
            t["all"] = (t["counter"] * t["all"] + probe_rat) / (t["counter"] + 1)
            t["sq"] = t["sq"] + probe_rat ** 2
            t["counter"] += 1

        def clamped(x):  # makeClamped, src/sampling.jl:99-106
            return np.clip(x, clamp[0], clamp[1]) if clamp is not None and len(clamp) else x

        sse = float(np.sum((t["vals"] - clamped(t["all"])) ** 2))
        sse_s = float(np.sum((t["vals"] - clamped(probe_rat)) ** 2))
        ok = float(np.sum((t["vals"] < t["cut"]) == (t["all"] < t["cut"])))
        return sse, sse_s, ok, float(len(t["vals"])), float(t["counter"])

    def get_test_predictions(self, rel, want_last=False):
        t = self._test[rel]
        return t["all"], t["sq"]

    def train_sse(self, rel):
        ents, ids, vals = self.rels[rel]
        err = orc.pred(ids, [self.U[k] for k in ents], self.mean[rel]) - vals
        return float(err @ err), len(vals)

    def sample_alpha(self, rel, lambda0, nu0, sse, n, chi2=float("nan")):
        self.calls.append(("alpha", rel))
        c2 = self.rng.chisquare(nu0 + n) if chi2 != chi2 else chi2
        self.alpha[rel] = (1.0 / (1.0 / lambda0 + sse)) * c2
        return self.alpha[rel]

    def advance_sweep(self):
        self.calls.append(("sweep",))

    def predict(self, rel, ids, test_F=None):
        ents, _, _ = self.rels[rel]
        return orc.pred(np.asarray(ids), [self.U[k] for k in ents], self.mean[rel])

    def predict_all(self, rel, shape):
        ents, _, _ = self.rels[rel]
        return self.U[ents[0]] @ self.U[ents[1]].T + self.mean[rel]

    # ---- entity side features (Macau): dense restatement, small sizes only ---------------------------------------------------
    def set_features(self, e, F):
        if hasattr(F, "rows") and hasattr(F, "cols"):
            Fd = np.zeros(F.shape)
            np.add.at(Fd, (np.asarray(F.rows) - 1, np.asarray(F.cols) - 1), 1.0)
        elif hasattr(F, "toarray"):
            Fd = F.toarray().astype(np.float64)
        else:
            Fd = np.asarray(F, dtype=np.float64)
        self.F = getattr(self, "F", {})
        self.beta = getattr(self, "beta", {})
        self.uhat = getattr(self, "uhat", {})
        self.FF = getattr(self, "FF", {})
        self.F[e], self.beta[e], self.uhat[e] = Fd, np.zeros((Fd.shape[1], self.D)), np.zeros((Fd.shape[0], self.D))

    def compute_ff(self, e, want=False):
        self.calls.append(("ff", e))
        self.FF[e] = self.F[e].T @ self.F[e]
        return self.FF[e] if want else None

    def update_uhat(self, e, mu, want=False):
        self.calls.append(("uhat", e))
        self.uhat[e] = self.F[e] @ self.beta[e]              # mj.uhat = F_mul_beta(en)' — src/macau.jl:103
        self._mu_rows = np.asarray(mu) + self.uhat[e]
        return self.uhat[e].copy() if want else None

    def sample_mode_uhat(self, e, Lambda, z=None):
        self.sample_mode(e, self._mu_rows, Lambda, z)

    def nw_stats_uhat(self, e):
        self.calls.append(("stats", e))
        self.stats[e] = orc.nw_stats(self.U[e], self.uhat[e])   # U = mj.sample - mj.uhat — src/macau.jl:124
        return self.stats[e]

    def beta_gram(self, e):
        return self.beta[e].T @ self.beta[e]

    def sample_beta(self, e, mu, Lambda, lambda_beta, tol=float("nan"), E1=None, E2=None, want_rhs=False, want_beta=True):
        self.calls.append(("beta", e, e in self.FF))
        F = self.F[e]
        N, numF = F.shape
        E1 = self.rng.standard_normal((N, self.D)) if E1 is None else E1
        E2 = self.rng.standard_normal((numF, self.D)) if E2 is None else E2
        rhs = F.T @ ((self.U[e] - np.asarray(mu)) + orc.color_noise(Lambda, E1)) + np.sqrt(lambda_beta) * orc.color_noise(Lambda, E2)
        self.beta[e] = orc.solve_full(F.T @ F, rhs, lambda_beta)   # CG and the FF solve have the same solution
        it = np.zeros(self.D, dtype=np.int32)
        return (self.beta[e].copy(), rhs, it) if want_rhs else (self.beta[e].copy(), it)

    def sample_lambda_beta(self, e, Lambda, nu, mu, gamma_variate=float("nan")):
        self.calls.append(("lambda_beta", e))
        numF = self.F[e].shape[1]
        g = self.rng.standard_gamma((nu + numF * self.D) / 2.0) if gamma_variate != gamma_variate else gamma_variate
        return orc.lambda_beta(orc.btb(self.beta[e]), Lambda, numF, nu, mu, g)

    def get_beta(self, e):
        return self.beta[e].copy()

    def close(self):
        pass
