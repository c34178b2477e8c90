'''
TEST INFRASTRUCTURE ONLY: per-record CPU restatement of `yanocomp prep`
(nanopolish eventalign TSV -> per-gene kmer-level event tables).

Only tests/ may import this.  It follows the reference record by record, with
plain Python loops, and returns the logical content the reference would write
to HDF5 (the writer itself needs h5py, which is not installable here):

  parse_eventalign_summary / parse_eventalign   reference yanocomp/io.py:41-85
  read_gtf / load_gtf_database                  reference yanocomp/io.py:88-143
  transcript_to_genomic                         reference yanocomp/prep.py:18-31
  calculate_kmer_level_stats                    reference yanocomp/prep.py:34-48
  collapse_chunk / get_chunks                   reference yanocomp/prep.py:51-130
  save_events (index assignment, binary16)      reference yanocomp/io.py:146-243

Parity status: the reference's tests hold no golden vector for `prep`; the pin
is the shipped fixture pair (tests/golden/*.hdf5): an eventalign table
regenerated from a fixture must collapse back to that fixture's events.
'''
import csv
import gzip
import itertools as it
import re
from bisect import bisect_right
from collections import OrderedDict

import numpy as np


def _lines(fn):
    if str(fn).endswith('.gz'):
        with gzip.open(fn, 'rt') as f:
            yield from f
    else:
        with open(fn) as f:
            yield from f


def parse_eventalign_summary(summary_fn):
    '''read_index -> read_name (io.py:41-48)'''
    return {r['read_index']: r['read_name'] for r in csv.DictReader(_lines(summary_fn), delimiter='\t')}


def parse_eventalign(eventalign_fn, summary_fn=None):
    '''io.py:51-85: position + 2, points = end_idx - start_idx'''
    names = parse_eventalign_summary(summary_fn) if summary_fn is not None else None
    parser = csv.DictReader(_lines(eventalign_fn), delimiter='\t')
    fields = set(parser.fieldnames)
    if 'read_name' not in fields and names is None:
        raise KeyError('read names missing')
    if not fields.issuperset(['start_idx', 'end_idx']):
        raise KeyError('signal index missing')
    for r in parser:
        yield {'t_id': r['contig'], 'pos': int(r['position']) + 2, 'kmer': r['reference_kmer'],
               'r_id': r['read_name'] if 'read_name' in r else names[r['read_index']],
               'mean': float(r['event_level_mean']), 'std': float(r['event_stdv']),
               'duration': float(r['event_length']), 'points': int(r['end_idx']) - int(r['start_idx'])}


def load_gtf_database(gtf_fn):
    '''io.py:98-143: exon records keyed by transcript_id; 0-based half-open intervals, cumulative offsets'''
    gtf = {}
    for line in _lines(gtf_fn):
        if line.startswith('#'):
            continue
        f = line.rstrip('\n').split('\t')
        if len(f) < 9 or f[2] != 'exon':
            continue
        attrs = dict(re.findall(r'(\w+) \"(.+?)\"(?:;|$)', f[8]))
        t = gtf.setdefault(attrs['transcript_id'], {'gene_id': attrs['gene_id'], 'chrom': f[0], 'strand': f[6], 'invs': []})
        t['invs'].append((int(f[3]) - 1, int(f[4])))
    for t in gtf.values():
        t['invs'].sort()
        lns = [e - s for s, e in t['invs']]
        t['ln'] = sum(lns)
        t['inv_cs'] = [0] + list(np.cumsum(lns)[:-1])
    return gtf


def transcript_to_genomic(pos, m):
    '''prep.py:18-31'''
    if not 0 <= pos < m['ln']:
        raise ValueError('position out of range')
    if m['strand'] == '-':
        pos = m['ln'] - pos - 1
    i = bisect_right(m['inv_cs'], pos) - 1
    return m['chrom'], m['invs'][i][0] + (pos - m['inv_cs'][i]), m['strand'], m['gene_id']


def calculate_kmer_level_stats(means, stdvs, ns):
    '''prep.py:34-48: pooled mean / std of several events of one kmer, weights = signal points'''
    var = stdvs ** 2
    pooled_var = sum(var * ns) / sum(ns)
    corr = 0
    for i, j in it.combinations(np.arange(len(var)), r=2):
        corr += ns[i] * ns[j] * (means[i] - means[j]) ** 2
    corr /= sum(ns) ** 2
    return sum(means * ns) / sum(ns), np.sqrt(pooled_var + corr)


def collapse(eventalign_fn, summary_fn=None, gtf_fn=None, chunksize=1_000_000):
    '''
    prep.py:51-148 + io.py:174-243 in one process and in file order.  Returns
    {gene_id: {'chrom','strand','events': [(t_idx, r_idx, pos, mean, std, duration) as binary16-rounded
    floats], 'kmers': {pos: kmer}, 'read_ids': [...], 'transcript_ids': [...]}}.
    '''
    gtf = load_gtf_database(gtf_fn) if gtf_fn is not None else None
    genes = OrderedDict()

    def flush(chunk):
        rows, info = [], {}
        key = lambda r: (r['t_id'], r['pos'], r['kmer'], r['r_id'])
        for (t_id, tpos, kmer, r_id), recs in it.groupby(chunk, key):
            if gtf is not None:
                chrom, gpos, strand, gene_id = transcript_to_genomic(tpos, gtf[t_id])
            else:
                chrom, gpos, strand, gene_id = '.', tpos, '.', t_id
            info[gene_id] = (chrom, strand)
            recs = list(recs)
            if len(recs) == 1:
                m, s = recs[0]['mean'], recs[0]['std']
            else:
                m, s = calculate_kmer_level_stats(np.asarray([r['mean'] for r in recs]),
                                                  np.asarray([r['std'] for r in recs]),
                                                  np.asarray([r['points'] for r in recs]))
            dur = 0
            for r in recs:
                dur += r['duration']
            rows.append((gene_id, t_id, r_id, gpos, kmer, m, s, dur))
        rows.sort(key=lambda r: r[0])
        for gene_id, grp in it.groupby(rows, lambda r: r[0]):
            g = genes.setdefault(gene_id, {'events': [], 'kmers': {}, 'read_ids': OrderedDict(),
                                           'transcript_ids': OrderedDict()})
            g['chrom'], g['strand'] = info[gene_id]
            for _, t_id, r_id, pos, kmer, m, s, dur in grp:
                t_x = g['transcript_ids'].setdefault(t_id, len(g['transcript_ids']))
                r_x = g['read_ids'].setdefault(r_id, len(g['read_ids']))
                with np.errstate(over='ignore'):
                    g['events'].append((t_x, r_x, pos, float(np.float16(m)), float(np.float16(s)),
                                        float(np.float16(dur))))
                g['kmers'][pos] = kmer

    chunk, curr = [], None
    for rec in parse_eventalign(eventalign_fn, summary_fn):          # prep.py:110-130
        if rec['r_id'] != curr and len(chunk) >= chunksize:
            flush(chunk)
            chunk = []
        curr = rec['r_id']
        chunk.append(rec)
    flush(chunk)
    for g in genes.values():
        g['read_ids'] = list(g['read_ids'])
        g['transcript_ids'] = list(g['transcript_ids'])
    return genes
