#!/usr/bin/env python
"""Drop-in for the reference's driver (/root/reference/pyMCZ/mcz.py): same command line, same
``run()`` signature, same input files, same table / ascii / pickle products -- with the
Monte-Carlo loop (sampling, metallicity.calculation, percentile extraction) executed by ONE
fused CUDA launch over all measurements (``pymcz_b200.ops.forward_production``) instead of a
Python loop over galaxies or a multiprocessing Pool.

What is NOT here (out of scope, host-side product output): matplotlib
histograms / boxplots, the Knuth / Doane / Bayesian-block binning and KDE that only feed those
plots (mcz.py:90-116, 227-239, 311-423, 678-703).  ``--noplot`` is therefore implied.
"""
from __future__ import print_function

import argparse
import csv
import os
import pickle
import sys

import numpy as np

from . import _scales as sc
from . import metallicity

__version__ = '1.3.1-b200'

NM0 = 0
alllines = list(sc.LINES)                                                  # mcz.py:45
morelines = ['E(B-V)', 'dE(B-V)', 'scale_blue', 'd scale_blue']            # mcz.py:46
MAXPROCESSES = 10                                                          # mcz.py:48 (CPU pool size of the reference)

CLOBBER = False
VERBOSE = False
UNPICKLE = False
ASCIIOUTPUT = False
ASCIIDISTRIB = False
RUNSIM = True
NOPLOT = True
BINMODE = 'k'

#: filled by run(): dict(res=..., pct=..., nvalid=..., status=..., trips=...) of the last call
last_result = None


def is_number(s):                                                          # mcz.py:72-79
    if not type(s) is np.bytes_:
        try:
            float(s)
            return True
        except ValueError:
            return False
    return False


def smart_open(logfile=None):                                              # mcz.py:82-87
    if logfile and logfile != '-':
        return open(logfile, 'w')
    return sys.stdout


def readfile(filename):
    """mcz.py:126-166: header-optional whitespace ascii -> structured array, fluxes as float32."""
    noheader = 1
    findex = -1
    with open(filename, 'r') as f:
        l0 = f.readline().replace(' ', '')
        l1 = f.readline().split()
    if l0.startswith('#') or l0.startswith(';'):
        header = l0.strip().replace(";", '').replace("#", '').split(',')
        header[0] = header[0].replace(' ', '')
        header = header[:len(l1)]
    else:
        noheader = 0
        header = ['galnum'] + alllines + ['flag'] + morelines
        header = header[:len(l1)]
    usecols, newheader = [], []
    for i, h in enumerate(header):
        if h in alllines + ['galnum']:
            usecols.append(i)
            newheader.append(h)
    header = newheader
    formats = ['S10'] + ['f'] * (len(header) - 1)
    if 'flag' in header:
        findex = header.index('flag')
        formats[findex] = 'S10'
    bstruct = {}
    for i, k in enumerate(header):
        bstruct[k] = [i, 0]
    b = np.loadtxt(filename, skiprows=noheader, dtype={'names': header, 'formats': formats},
                   comments=';', usecols=usecols)
    if b.size == 1:
        b = np.atleast_1d(b)
    for i, k in enumerate(header):
        if not k == 'flag' and is_number(b[k][0]):
            bstruct[k][1] = np.count_nonzero(b[k]) + sum(np.isnan(b[k]))
    j = len(b['galnum'])
    return b, j, bstruct


def ingest_data(filename, path):                                           # mcz.py:169-191
    measfile = os.path.join(path, filename + "_meas.txt")
    errfile = os.path.join(path, filename + "_err.txt")
    meas, nm, bsmeas = readfile(measfile)
    err, nm, bserr = readfile(errfile)
    # the reference's SNR<3 prompt (mcz.py:185-188) is dead code (`.any() < 3`); a library must
    # not block on stdin anyway
    return (filename, meas, err, nm, path, (bsmeas, bserr))


def input_data(filename, path):                                            # mcz.py:193-201
    p = os.path.join(path, "input")
    assert os.path.isdir(p), "bad data directory %s" % p
    if os.path.isfile(os.path.join(p, filename + '_err.txt')):
        if os.path.isfile(os.path.join(p, filename + '_meas.txt')):
            return ingest_data(filename, path=p)
    print("Unable to find _meas and _err files ", filename + '_meas.txt', filename + '_err.txt', "in directory ", p)
    return -1


def errordistrib(distrargs, n, distype='normal'):                          # mcz.py:207-217
    if distype == 'normal':
        try:
            mu, sigma = distrargs
        except Exception:
            print("for normal distribution distargs must be a 2 element tuple")
            return -1
        return np.random.normal(mu, sigma, n)
    print("distribution not supported")
    return -1


def tables_from_structured(flux, err, nm):
    """Structured arrays of readfile() -> float32 [nm, 11] tables in ``alllines`` order; a line
    that is not a column of the file is NaN (absent) with zero error."""
    meas_t = np.full((nm, sc.NLINES), np.nan, dtype=np.float32)
    err_t = np.zeros((nm, sc.NLINES), dtype=np.float32)
    for j, name in enumerate(alllines):
        if name in flux.dtype.names:
            meas_t[:, j] = flux[name]
        if name in err.dtype.names:
            err_t[:, j] = err[name]
    return meas_t, err_t


def savehist_line(snname, key, status, median, pc16, pc84, reserr=None):
    """The statistics / printing part of savehist() (mcz.py:273-306, 423) from the percentile
    table the GPU returned.  Returns the tab-separated string savehist returns."""
    if status == sc.ST_EMPTY:                                              # mcz.py:277-280
        return "-1\t -1\t -1"
    if status == sc.ST_NONPOS:                                             # mcz.py:282-285
        print('{0:15} {1:20} {2:>13d}   {3:>7d}   {4:>7d} '.format(snname, key, -1, -1, -1))
        return "-1\t -1\t -1"
    median, left, right = float(median), float(pc16), float(pc84)
    if round(right, 6) == round(left, 6) and round(left, 6) == round(median, 6):      # mcz.py:297-301
        print('{0:15} {1:20} {2:>13.3f}   -{3:>7.3f}   +{4:>7.3f} (no distribution)'.format(snname, key, median, 0, 0))
    else:
        print('{0:15} {1:20} {2:>13.3f}   -{3:>7.3f}   +{4:>7.3f}'.format(
            snname, key, round(median, 3), round(median - left, 3), round(right - median, 3)))   # mcz.py:303-306
    if reserr:
        print('+/- {0:.3f}'.format(reserr))
    return "%f\t %f\t %f" % (round(median, 3), round(median - left, 3), round(right - median, 3))


def run(fi, nsample, mds, multiproc, logf, unpickle=False, dust_corr=True, verbose=False, fs=24,
        seed=None, keep_samples=None):
    """mcz.py:470-709.  ``fi`` is the tuple input_data() returns.

    ``multiproc`` selects the reference's Pool semantics -- one deviate per sample shared by all
    lines (mcz.py:442) -- otherwise lines are perturbed independently like the serial loop
    (mcz.py:580-589).  ``seed`` (extra, optional) fixes the Philox key; the reference is unseeded.
    ``keep_samples`` (extra): keep the per-sample distributions for the pickle / --asciidistrib
    products; default: only when they fit comfortably (nm * 37 * N' * 4 bytes <= 2 GiB).
    """
    global RUNSIM, last_result
    import torch
    from . import dist

    (name, flux, err, nm, path, bss) = fi
    assert (len(flux[0]) == len(err[0])), "flux and err must be same dimensions"
    assert (len(flux['galnum']) == nm), "flux and err must be of declaired size"
    assert (len(err['galnum']) == nm), "flux and err must be same dimensions"

    newnsample = nsample                                                   # mcz.py:476-479
    if nsample > 1:
        newnsample = int(nsample + 0.1 * nsample)
    p = os.path.join(path, '..')
    Zs = metallicity.get_keys()
    binp = os.path.join(p, 'output', '%s' % name)
    os.makedirs(os.path.join(binp, 'hist'), exist_ok=True)                 # mcz.py:488-491
    picklefile = os.path.join(binp, '%s_n%d.pkl' % (name, nsample))
    if VERBOSE:
        print("output files will be stored in ", binp)

    res = None
    RUNSIM = True
    if unpickle:                                                           # mcz.py:500-507
        if os.path.isfile(picklefile):
            with open(picklefile, 'rb') as pklfile:
                res = pickle.load(pklfile)
            RUNSIM = False
        else:
            print("missing pickled file for this simulation: running the MonteCarlo")

    meas_t, err_t = tables_from_structured(flux, err, nm)
    tokens, third, _unknown = sc.parse_md(mds)
    if metallicity.DROPNEGATIVES:                                          # metallicity.py:24, 101-102
        tokens |= sc.T_DROPNEG
    for t in third + (['D13'] if 'all' in mds.split(',') else []):
        metallicity.printsafemulti("WARNING: %s scales need third-party code (pyqz / HII-CHI-mistry) that this "
                                   "build does not vendor; they are left empty" % t, logf, 1)
    if keep_samples is None:
        keep_samples = nm * sc.NKEYS * newnsample * 4 <= (2 << 30)

    if RUNSIM:
        # The per-measurement log of the reference's serial loop (mcz.py:577-590: header, the measured
        # fluxes, then the "calculating X" line of every scale calculation() dispatches, metscales.py:388-1196).
        # The Pool path logs nothing per measurement; neither do tables too long for a text log.
        if not multiproc and dist.rank() == 0 and logf is not None and nm <= 100000:
            names = metallicity.dispatch_order(mds.split(','))
            for i in range(NM0, nm):
                logf.write("\n\n measurements %d\n" % (i + 1))
                for k in alllines:
                    if k in flux.dtype.names:
                        logf.write('{0:15} '.format(k))
                        logf.write('{0:0.2} +/- {1:0.2}\n'.format(flux[k][i], err[k][i]))
                logf.write("\n")
                if dust_corr and np.isfinite(flux['Ha'][i]) and np.isfinite(flux['Hb'][i]):
                    logf.write("calculating E(B-V)\n")
                logf.write("calculating R23\n")
                for nme in names:
                    logf.write("calculating %s\n" % nme)
        if seed is None:
            seed = int.from_bytes(os.urandom(8), 'little')                 # the reference is unseeded
        if dist.world_size() > 1:                                          # one realisation: rank 0's seed everywhere
            import torch.distributed as td
            box = [seed]
            td.broadcast_object_list(box, src=0)
            seed = int(box[0])
        out = dist.production_table(meas_t, err_t, newnsample, tokens,
                                    sc.DUST_ON if dust_corr else sc.DUST_NONE, seed,
                                    correlated=bool(multiproc), deterministic=(nsample == 1),
                                    return_samples=keep_samples)
        pct, nvalid, status, trips, ret = out["pct"], out["nvalid"], out["status"], out["trips"], out["ret"]
        for i in np.nonzero(ret == -1)[0]:                                 # mcz.py:595-596
            print("measurement %d: MINIMUM REQUIRED LINES:  [OII]3727 & [OIII]5007, or [NII]6584, "
                  "and Ha & Hb if you want dereddening" % (i + 1))
        if out["Z"] is not None:
            # res[key] -> (N', nm) float64, NaN where the scale is None (mcz.py:600-611)
            res = {key: np.ascontiguousarray(out["Z"][:, j, :].T.astype(np.float64)) for j, key in enumerate(Zs)}
            if dist.rank() == 0:
                with open(picklefile, 'wb') as fpk:                        # mcz.py:618
                    pickle.dump(res, fpk)
        last_result = dict(res=res, pct=pct, nvalid=nvalid, status=status, trips=trips, ret=ret, seed=seed)
    else:
        # statistics of a pickled realisation: the float64 percentile op on the stored samples
        from . import ops
        Zt = np.stack([np.asarray(res[key], dtype=np.float64).T for key in Zs], axis=1)   # [nm, 37, N']
        present = ~np.all(np.isnan(Zt), axis=2)
        pct_t, nvalid_t, status_t = ops.segmented_percentiles(torch.from_numpy(np.ascontiguousarray(Zt)).cuda(),
                                                              torch.from_numpy(present).cuda())
        pct, nvalid, status = pct_t.cpu().numpy(), nvalid_t.cpu().numpy(), status_t.cpu().numpy()
        last_result = dict(res=res, pct=pct, nvalid=nvalid, status=status, trips=None, ret=None, seed=None)

    if dist.rank() != 0:
        return None

    # ---- the result table (mcz.py:627-676)
    print("\n\n")
    print('{0:15} {1:20} {2:>13}   -{3:>5}     +{4:>5}  {5:11} {6:>7}'
          .format("SN", "diagnostic", "metallicity", "34%", "34%", "(sample size:", '%d)' % nsample))
    for i in range(NM0, nm):
        fiout = None
        if ASCIIOUTPUT:
            fiout = open(os.path.join(binp, '%s_n%d_%d.txt' % (name, nsample, i + 1)), 'w')
            fiout.write("%s\t Median Oxygen abundance (12+log(O/H))\t 16th percentile\t 84th percentile\n" % name)
        galname = flux[i]['galnum']
        galname = galname.decode() if isinstance(galname, bytes) else galname
        print("\n\nmeasurement ",
              "%d/%d : %s-------------------------------------------------------------" % (i + 1, nm, galname))
        for j, key in enumerate(Zs):
            if status[i, j] == sc.ST_ABSENT or status[i, j] == sc.ST_EMPTY:      # sum(~isnan) > 0, mcz.py:651-652
                continue
            if ASCIIDISTRIB and res is not None:
                with open(os.path.join(binp, '%s_n%d_%s_%d.csv' % (name, nsample, key, i + 1)), "w", newline='') as fidist:
                    csv.writer(fidist).writerow(res[key][:, i])
            sh = savehist_line(name, key, status[i, j], pct[i, j, 0], pct[i, j, 1], pct[i, j, 2])
            if fiout is not None:
                fiout.write(key + "\t " + sh + '\n')
        if fiout is not None:
            fiout.close()
        if VERBOSE:
            print("uncertainty calculation complete")
    return None


def main(argv=None):
    """mcz.py:711-770: identical flags."""
    parser = argparse.ArgumentParser()
    parser.add_argument('name', metavar='<name>', type=str, help="the SN file name (root of the _min,_max file names")
    parser.add_argument('nsample', metavar='N', type=int, help="number of iterations, minimum 100 (or 1 for no MC sampling)")
    parser.add_argument('--clobber', default=False, action='store_true', help="replace existing output")
    parser.add_argument('--binmode', default='k', type=str, choices=['d', 's', 'k', 't', 'bb', 'kd'],
                        help="histogram bin rule (kept for compatibility: this build draws no histograms)")
    parser.add_argument('--path', default=None, type=str,
                        help="input/output path (must contain the input _meas.txt and _err.txt files in a subdirectory input)")
    parser.add_argument('--unpickle', default=False, action='store_true', help="read the pickled realization instead of making a new one")
    parser.add_argument('--verbose', default=False, action='store_true', help="verbose mode")
    parser.add_argument('--log', default=None, type=str, help="log file, if not passed defaults to standard output")
    parser.add_argument('--nodust', default=False, action='store_true', help=" don't do dust corrections (default is to do it)")
    parser.add_argument('--noplot', default=False, action='store_true', help=" don't plot individual distributions (always the case in this build)")
    parser.add_argument('--asciiout', default=False, action='store_true', help=" write distribution in an ascii output (default is not to)")
    parser.add_argument('--asciidistrib', default=False, action='store_true', help=" write distribution in an ascii output (default is not to)")
    parser.add_argument('--md', default='all', type=str,
                        help="metallicity diagnostic to calculate. default is 'all', options are: D02, Z94, M91, C01, P05, "
                             "M08, M08all, M13, PP04, D13, D13all, KD02, DP00 (deprecated), P01, D16")
    parser.add_argument('--multiproc', default=False, action='store_true',
                        help=" the reference's Pool semantics: one deviate per sample shared by all lines")
    parser.add_argument('--seed', default=None, type=int, help="(extra) Philox seed; default: from the OS")
    args = parser.parse_args(argv)

    global CLOBBER, VERBOSE, BINMODE, ASCIIOUTPUT, ASCIIDISTRIB, NOPLOT
    CLOBBER = args.clobber
    VERBOSE = args.verbose
    BINMODE = args.binmode
    NOPLOT = True
    ASCIIOUTPUT = args.asciiout
    ASCIIDISTRIB = args.asciidistrib

    if args.path:
        path = args.path
    else:
        assert (os.getenv("MCMetdata")), "pass --path or set the environmental variable MCMetdata"
        path = os.getenv("MCMetdata")
    assert (os.path.isdir(path)), "pass a path or set up the environmental variable MCMetdata"
    if args.nsample == 1:
        print("CALCULATING METALLICITY WITHOUT GENERATING MC DISTRIBUTIONS")
    if args.nsample == 1 or args.nsample >= 10:                             # mcz.py:762
        fi = input_data(args.name, path=path)
        if fi != -1:
            logf = smart_open(args.log)
            run(fi, args.nsample, args.md, args.multiproc, logf, unpickle=args.unpickle,
                dust_corr=(not args.nodust), verbose=VERBOSE, seed=args.seed)
            if args.log:
                logf.close()
    else:
        print("nsample must be at least 100")


if __name__ == "__main__":
    main()
