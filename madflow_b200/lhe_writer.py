"""Les Houches Event output -- python_package/madflow/lhe_writer.py.

Same surface as the reference's writer (`LheWriter(folder, run, no_unweight, event_target)` as a context
manager, `lhe_parser(all_ps, res)`, `store_result`, `dump_result`, `cross`, `err`, `do_unweighting`,
`EventFileFlow`, `FourMomentumFlow`), same files (`<folder>/Events/<run>/weighted_events.lhe.gz`,
`unweighted_events.lhe.gz`, lhe_writer.py:113-121, 334-346).

Differences, all forced by what is available offline:
  * the reference builds every event with MG5_aMC's `madgraph.various.lhe_parser` (imported at module load,
    lhe_writer.py:18-27); MG5 is not part of the reference tree, so the <event> blocks are formatted here, in
    the layout MG5's `Event.__str__` / `Particle.__str__` produce ("parity unpinned");
  * the reference hard-wires the particle ids of p p > t t~ (lhe_writer.py:173-185); here they come from the
    process (`pdg`, default g g > t t~ + gluons by multiplicity);
  * events arrive as arrays (numpy / torch, host or device), not through tf.py_function.
"""
import gzip
import logging
import math
from multiprocessing.pool import ThreadPool as Pool
from pathlib import Path
from time import time as tm

import numpy as np

logger = logging.getLogger("madflow")


def _to_numpy(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def default_pdg(nexternal):
    """g g > t t~ (+ gluons): the processes this package generates itself."""
    return [21, 21, 6, -6] + [21] * (nexternal - 4)


def format_event(ps, wgt, pdg, aqcd=0.0, scale=0.0, aqed=0.0, ievent=1):
    """One <event> block.  ps: (nexternal, 4) rows (E, px, py, pz); the first two rows are the incoming
    particles (status -1), the others outgoing (status 1, mothers 1 2).  Colour lines are 0 like in the
    reference (lhe_writer.py:187-204)."""
    n = len(ps)
    lines = ["<event>", "%2d %6d %+13.7e %14.8e %14.8e %14.8e" % (n, ievent, wgt, scale, aqed, aqcd)]
    for i, (e, px, py, pz) in enumerate(ps):
        m2 = e * e - px * px - py * py - pz * pz
        mass = math.sqrt(m2) if m2 > 0.0 else 0.0
        status, m1, m2_ = (-1, 0, 0) if i < 2 else (1, 1, 2)
        lines.append("%8d %2d %4d %4d %4d %4d %+13.10e %+13.10e %+13.10e %14.10e %14.10e %10.4e %10.4e"
                     % (pdg[i], status, m1, m2_, 0, 0, px, py, pz, e, mass, 0.0, 0.0))
    lines.append("</event>")
    return "\n".join(lines) + "\n"


class FourMomentumFlow:
    """(E, px, py, pz) with the kinematic properties used by example/compare_mg5_hists.py
    (lhe_writer.py:385-433 on top of MG5's FourMomentum)."""

    def __init__(self, obj=0, px=0, py=0, pz=0, E=0):
        if isinstance(obj, ParticleFlow) or isinstance(obj, FourMomentumFlow):
            px, py, pz, E = obj.px, obj.py, obj.pz, obj.E
        elif isinstance(obj, (list, tuple)):
            E, px, py, pz = obj
        elif isinstance(obj, str):
            E, px, py, pz = map(float, obj.split())
        elif obj:
            E = obj
        self.E, self.px, self.py, self.pz = float(E), float(px), float(py), float(pz)

    @property
    def pt(self):
        return math.sqrt(self.px**2 + self.py**2)

    @property
    def pseudorapidity(self):
        norm = math.sqrt(self.px**2 + self.py**2 + self.pz**2)
        return 0.5 * math.log((norm + self.pz) / (norm - self.pz))

    @property
    def rapidity(self):
        return 0.5 * math.log((self.E + self.pz) / (self.E - self.pz))

    @property
    def mass(self):
        m2 = self.E**2 - self.px**2 - self.py**2 - self.pz**2
        return math.sqrt(m2) if m2 > 0 else 0.0

    @property
    def phi(self):
        return 0.0 if self.pt == 0.0 else math.atan2(self.py, self.px)


class ParticleFlow:
    FIELDS = ("pid", "status", "mother1", "mother2", "color1", "color2", "px", "py", "pz", "E", "mass", "vtim",
              "helicity")

    def __init__(self, info):
        for f in self.FIELDS:
            setattr(self, f, info.get(f))


class EventFlow(list):
    """One event: header fields + a list of ParticleFlow (lhe_writer.py:32-67)."""

    def __init__(self, info):
        super().__init__()
        for f in ("nexternal", "ievent", "wgt", "aqcd", "scale", "aqed", "tag", "comment"):
            setattr(self, f, info.get(f))

    def add_particles(self, particles):
        self.extend(particles)

    def __str__(self):
        ps = [(p.E, p.px, p.py, p.pz) for p in self]
        return format_event(ps, self.wgt, [p.pid for p in self], self.aqcd or 0.0, self.scale or 0.0, self.aqed or 0.0,
                            self.ievent or 1)

    def as_bytes(self):
        return str(self).encode("utf-8")


class EventFileFlow:
    """Iterates over the <event> blocks of an .lhe / .lhe.gz file (lhe_writer.py:360-383)."""

    def __init__(self, path, mode="r"):
        self.path = Path(path)

    def _open(self):
        return gzip.open(self.path, "rt") if self.path.suffix == ".gz" else open(self.path, "r")

    def __iter__(self):
        with self._open() as fh:
            block = None
            for line in fh:
                s = line.strip()
                if s.startswith("<event"):
                    block = []
                elif s.startswith("</event"):
                    yield self._parse(block)
                    block = None
                elif block is not None and s:
                    block.append(s)

    @staticmethod
    def _parse(block):
        h = block[0].split()
        evt = EventFlow({"nexternal": int(h[0]), "ievent": int(h[1]), "wgt": float(h[2]), "scale": float(h[3]),
                         "aqed": float(h[4]), "aqcd": float(h[5]), "tag": "", "comment": ""})
        parts = []
        for row in block[1:1 + evt.nexternal]:
            f = row.split()
            vals = [int(v) for v in f[:6]] + [float(v) for v in f[6:13]]
            parts.append(ParticleFlow(dict(zip(ParticleFlow.FIELDS, vals))))
        evt.add_particles(parts)
        return evt

    def __len__(self):
        return sum(1 for _ in self)

    def unweight(self, outpath, event_target=0, seed=1234):
        """Hit-or-miss on |wgt| / max|wgt| (what MG5's EventFile.unweight does to first order); the kept events
        are written to `outpath` with their sign and unit modulus.  Returns the number kept."""
        wmax = max((abs(e.wgt) for e in self), default=0.0)
        rng = np.random.default_rng(seed)
        kept = 0
        with gzip.open(outpath, "wb") as out:
            out.write(b"<LesHouchesEvent>\n")
            for e in self:
                if wmax > 0.0 and abs(e.wgt) >= rng.random() * wmax:
                    e.wgt = math.copysign(1.0, e.wgt)
                    out.write(e.as_bytes())
                    kept += 1
                    if event_target and kept >= event_target:
                        break
            out.write(b"</LesHouchesEvent>\n")
        return kept


class LheWriter:
    def __init__(self, folder, run="run_01", no_unweight=False, event_target=0, pdg=None):
        """Writes LHE events to <folder>/Events/<run>/weighted_events.lhe.gz (lhe_writer.py:94-121)."""
        self.folder = Path(folder)
        self.run = run
        self.no_unweight = no_unweight
        self.event_target = event_target
        self.pdg = list(pdg) if pdg is not None else None
        self.pool = Pool(processes=1)
        lhe_folder = self.folder.joinpath(f"Events/{self.run}")
        lhe_folder.mkdir(parents=True, exist_ok=True)
        self.lhe_path = lhe_folder.joinpath("weighted_events.lhe.gz")
        self.stream = gzip.open(self.lhe_path, "wb")
        self.__cross = self.__err = None
        self.nevents = 0

    def __enter__(self):
        self.dump_banner()
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        """Close the asynchronous dumping pool; unweight unless no_unweight (lhe_writer.py:129-149).
        store_result() must have been called before when unweighting."""
        self.pool.close()
        self.pool.join()
        self.dump_exit()
        self.stream.close()
        logger.debug("Saved LHE file at %s", self.lhe_path.as_posix())
        if not self.no_unweight and exc_type is None:
            start = tm()
            nb_keep, nb_wgt = self.do_unweighting(event_target=self.event_target)
            logger.info("Unweighting stats: kept %d events out of %d (efficiency %.2g %%, time %.5f)",
                        nb_keep, nb_wgt, nb_keep / max(nb_wgt, 1) * 100, tm() - start)

    def lhe_parser(self, all_ps, res, alpha_s=None):
        """all_ps (nevents, nexternal, 4), res (nevents,) weights -- the two arguments the reference's
        integrand passes (lhe_writer.py:151-208, madflow_exec.py:462-464).  Events of weight 0 (cut away) are
        not written."""
        ps, wgt = _to_numpy(all_ps), _to_numpy(res).reshape(-1)
        aq = _to_numpy(alpha_s).reshape(-1) if alpha_s is not None else None
        keep = np.nonzero(wgt)[0]
        self.dump(ps[keep], wgt[keep], aq[keep] if aq is not None else None)
        return 0.0

    def dump_banner(self, stream=None):
        (stream or self.stream).write(b"<LesHouchesEvent>\n")

    def dump_exit(self, stream=None):
        (stream or self.stream).write(b"</LesHouchesEvent>\n")

    def dump_events(self, ps, wgt, aqcd=None):
        pdg = self.pdg or default_pdg(ps.shape[1])
        chunks = [format_event(p, w, pdg, aqcd=(aqcd[i] if aqcd is not None else 0.0))
                  for i, (p, w) in enumerate(zip(ps.tolist(), wgt.tolist()))]
        self.stream.write("".join(chunks).encode("utf-8"))
        self.nevents += len(chunks)

    def async_dump(self, *args):
        self.dump_events(*args)

    def dump(self, *args):
        """Dumps asynchronously (one worker thread, order preserved), lhe_writer.py:266-268."""
        self.pool.apply_async(self.async_dump, args)

    def dump_result(self, filename):
        """cross section and statistical error -> text file (lhe_writer.py:270-280)."""
        np.savetxt(Path(filename).as_posix(), np.array([self.__cross, self.__err]))

    @property
    def cross(self):
        return self.__cross

    @cross.setter
    def cross(self, value):
        self.__cross = value

    @property
    def err(self):
        return self.__err

    @err.setter
    def err(self, value):
        self.__err = value

    def store_result(self, result):
        self.__cross = float(result[0])
        self.__err = float(result[1])

    def do_unweighting(self, event_target=0):
        """weighted_events.lhe.gz -> unweighted_events.lhe.gz, every kept event with wgt = cross section
        (lhe_writer.py:316-356)."""
        lhe = EventFileFlow(self.lhe_path)
        nb_wgt = len(lhe)
        tmp_path = self.lhe_path.with_name("tmp_unweighted_events.lhe.gz")
        nb_keep = lhe.unweight(tmp_path.as_posix(), event_target=event_target)
        unwgt_path = tmp_path.with_name("unweighted_events.lhe.gz")
        with gzip.open(unwgt_path, "wb") as stream:
            self.dump_banner(stream)
            for event in EventFileFlow(tmp_path):
                event.wgt = math.copysign(self.__cross if self.__cross is not None else 1.0, event.wgt)
                stream.write(event.as_bytes())
            self.dump_exit(stream)
        tmp_path.unlink()
        return nb_keep, nb_wgt
