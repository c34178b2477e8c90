'''
`yanocomp prep`: nanopolish eventalign TSV -> kmer-level event tables
(reference yanocomp/prep.py).

Same command, options and per-gene logical output as the reference, with two
deliberate differences:

* records are handled as columns, not as one dict per line: the TSV is parsed
  by Arrow's CSV reader (`pyarrow.csv`; its number conversion is correctly
  rounded, so every value is the same double Python's float() gives; pyarrow
  is required by this command and its absence is reported when it starts), runs
  of events of one
  (transcript, position, kmer, read) are found with array comparisons, and
  the pooled statistics are evaluated for all runs of the same length at once
  IN THE REFERENCE'S ORDER OF OPERATIONS (left-to-right sums, the i<j pair
  loop of prep.py:41-44), so the binary16 values written agree with a
  per-record evaluation;
* the output is the packed container of `packed.py` (gmmtest reads it
  directly; no HDF5 library is available offline to write the reference's
  format).  `-h/--hdf5-fn` keeps its name so command lines carry over.

`-p/--processes` is accepted for compatibility; the collapse is vectorised in
one process (the reference needs the pool because it loops per record).
'''
import itertools as it
import logging
import os
from bisect import bisect_right

import click
import numpy as np

from .io import parse_eventalign, load_gtf_database, save_events_to_packed

logger = logging.getLogger('yanocomp')


def transcript_to_genomic(pos, genomic_mapping):
    '''transcript coordinate -> (chrom, genomic pos, strand, gene_id) (reference prep.py:18-31)'''
    ln = genomic_mapping['ln']
    if not 0 <= pos < ln:
        raise ValueError('position out of range')
    if genomic_mapping['strand'] == '-':
        pos = ln - pos - 1
    inv_cs = genomic_mapping['inv_cs']
    exon_idx = bisect_right(inv_cs, pos) - 1
    return (genomic_mapping['chrom'], genomic_mapping['invs'][exon_idx][0] + (pos - inv_cs[exon_idx]),
            genomic_mapping['strand'], genomic_mapping['gene_id'])


def _transcripts_to_genomic(pos, genomic_mapping):
    '''the same for an array of positions of one transcript'''
    ln = genomic_mapping['ln']
    if len(pos) and (pos.min() < 0 or pos.max() >= ln):
        raise ValueError('position out of range')
    if genomic_mapping['strand'] == '-':
        pos = ln - pos - 1
    inv_cs = np.asarray(genomic_mapping['inv_cs'], dtype=np.int64)
    starts = np.asarray([s for s, _ in genomic_mapping['invs']], dtype=np.int64)
    exon_idx = np.searchsorted(inv_cs, pos, side='right') - 1
    return starts[exon_idx] + (pos - inv_cs[exon_idx])


def _seq_sum(m):
    '''row sums of a (G, k) matrix added left to right, as Python's sum() over an array does'''
    s = m[:, 0].copy()
    for c in range(1, m.shape[1]):
        s = s + m[:, c]
    return s


def calculate_kmer_level_stats(means, stdvs, ns):
    '''
    Pooled mean and std of the events of one kmer, weights = signal points
    (reference prep.py:34-48).  Accepts the reference's 1-D arrays (one kmer)
    or (G, k) matrices (G kmers of k events each).
    '''
    one = np.ndim(means) == 1
    means, stdvs = np.atleast_2d(np.asarray(means, dtype=np.float64)), np.atleast_2d(np.asarray(stdvs, dtype=np.float64))
    ns = np.atleast_2d(np.asarray(ns, dtype=np.int64))
    var = stdvs * stdvs
    n_tot = ns.sum(axis=1)
    with np.errstate(divide='ignore', invalid='ignore'):
        pooled_var = _seq_sum(var * ns) / n_tot
        corr = np.zeros(len(means))
        for i, j in it.combinations(range(means.shape[1]), r=2):
            d = means[:, i] - means[:, j]
            corr = corr + (ns[:, i] * ns[:, j]) * (d * d)
        corr = corr / (n_tot ** 2)
        pooled_std = np.sqrt(pooled_var + corr)
        pooled_mean = _seq_sum(means * ns) / n_tot
    return (pooled_mean[0], pooled_std[0]) if one else (pooled_mean, pooled_std)


def collapse_chunk(ea_chunk, to_gene=False):
    '''
    Collapse one chunk of eventalign columns to kmer level (reference
    prep.py:51-100).  `ea_chunk` = (columns, gtf or None) with columns =
    dict(t_code, t_ids, pos, k_code, kmers, r_code, r_ids, mean, std, duration, points).
    Returns (pieces, gene_info): pieces = list of (gene_id, columns of that gene's collapsed records in
    arrival order), gene_info[gene_id] = (chrom, strand).
    '''
    c, gtf = ea_chunk
    n = len(c['pos'])
    if n == 0:
        return [], {}
    t_code, pos, k_code, r_code = c['t_code'], c['pos'], c['k_code'], c['r_code']
    new = (t_code[1:] != t_code[:-1]) | (pos[1:] != pos[:-1]) | (k_code[1:] != k_code[:-1]) | (r_code[1:] != r_code[:-1])
    starts = np.concatenate(([0], np.flatnonzero(new) + 1))
    counts = np.diff(np.concatenate((starts, [n])))

    mean = c['mean'][starts].copy()
    std = c['std'][starts].copy()
    dur = c['duration'][starts].copy()
    for k in np.unique(counts):
        if k == 1:
            continue                                   # a single event is taken as it is (prep.py:83-85)
        sel = np.flatnonzero(counts == k)
        idx = starts[sel][:, None] + np.arange(k)
        mean[sel], std[sel] = calculate_kmer_level_stats(c['mean'][idx], c['std'][idx], c['points'][idx])
        dur[sel] = _seq_sum(c['duration'][idx])

    g_t, g_pos = t_code[starts], pos[starts]
    gene_info = {}
    if to_gene:
        gene_of_t = np.empty(len(c['t_ids']), dtype=np.int64)
        gene_names, gene_lookup = [], {}
        gpos = np.empty_like(g_pos)
        for tc in np.unique(g_t):
            m = gtf[c['t_ids'][tc]]                    # KeyError for a transcript the GTF lacks, as in the reference
            sel = np.flatnonzero(g_t == tc)
            gpos[sel] = _transcripts_to_genomic(g_pos[sel], m)
            gi = gene_lookup.setdefault(m['gene_id'], len(gene_names))
            if gi == len(gene_names):
                gene_names.append(m['gene_id'])
            gene_of_t[tc] = gi
            gene_info[m['gene_id']] = (m['chrom'], m['strand'])
        g_gene = gene_of_t[g_t]
    else:
        gene_names, g_gene, gpos = list(c['t_ids']), g_t, g_pos
        for tc in np.unique(g_t):
            gene_info[gene_names[tc]] = ('.', '.')

    order = np.argsort(g_gene, kind='stable')          # pieces per gene, arrival order kept (prep.py:98-99)
    cuts = np.flatnonzero(np.diff(g_gene[order])) + 1
    pieces = []
    for sel in np.split(order, cuts):
        pieces.append((gene_names[g_gene[sel[0]]], {
            't_code': g_t[sel], 't_ids': c['t_ids'], 'r_code': r_code[starts][sel], 'r_ids': c['r_ids'],
            'pos': gpos[sel], 'k_code': k_code[starts][sel], 'kmers': c['kmers'],
            'mean': mean[sel], 'std': std[sel], 'duration': dur[sel]}))
    return pieces, gene_info


def get_chunks(ea, approx_chunksize=None, gtf=None):
    '''
    Re-cut the parser's column blocks so that a chunk never ends inside a read
    (reference prep.py:110-130): the rows of the last read of each block are
    carried into the next one.
    '''
    carry = None
    for block in ea:
        if carry is not None:
            block = _concat_blocks(carry, block)
            carry = None
        r = block['r_code']
        change = np.flatnonzero(r[1:] != r[:-1])
        if len(change) == 0:
            carry = block
            continue
        cut = int(change[-1]) + 1
        carry = _slice_block(block, cut, len(r))
        yield _slice_block(block, 0, cut), gtf
    if carry is not None and len(carry['pos']):
        yield carry, gtf


_ROW_KEYS = ('t_code', 'pos', 'k_code', 'r_code', 'mean', 'std', 'duration', 'points')


def _slice_block(b, lo, hi):
    out = {k: b[k][lo:hi] for k in _ROW_KEYS}
    out.update(t_ids=b['t_ids'], kmers=b['kmers'], r_ids=b['r_ids'])
    return out


def _concat_blocks(a, b):
    '''append block b to block a, re-basing b's string codes on a's tables'''
    out = {}
    for code, table in (('t_code', 't_ids'), ('k_code', 'kmers'), ('r_code', 'r_ids')):
        names = list(a[table])
        lookup = {s: i for i, s in enumerate(names)}
        remap = np.empty(len(b[table]), dtype=np.int64)
        for i, s in enumerate(b[table]):
            j = lookup.get(s)
            if j is None:
                j = lookup[s] = len(names)
                names.append(s)
            remap[i] = j
        out[table] = names
        out[code] = np.concatenate((a[code], remap[b[code]] if len(b[code]) else b[code]))
    for k in ('pos', 'mean', 'std', 'duration', 'points'):
        out[k] = np.concatenate((a[k], b[k]))
    return out


def parallel_collapse(ea_fn, summary_fn=None, gtf_fn=None, processes=1, chunksize=1_000_000):
    '''yields (gene_id, chrom, strand, records) pieces (reference prep.py:133-165)'''
    if gtf_fn is not None:
        logger.info(f'Loading GTF records from {os.path.abspath(gtf_fn)}')
        gtf = load_gtf_database(gtf_fn, mapped_to='transcript_id', parent_id='gene_id')
        to_gene = True
        logger.info(f'{len(gtf):,} transcript records loaded from GTF')
    else:
        logger.info('No GTF provided, records will be grouped by transcript')
        gtf, to_gene = None, False
    for chunk in get_chunks(parse_eventalign(ea_fn, summary_fn, chunksize), chunksize, gtf):
        pieces, gene_info = collapse_chunk(chunk, to_gene=to_gene)
        for gene_id, records in pieces:
            chrom, strand = gene_info[gene_id]
            yield gene_id, chrom, strand, records


class CollapseOpts:
    def __init__(self, **kw):
        self.__dict__.update(kw)


@click.command(short_help='Convert nanopolish eventalign tsv to a packed event file')
@click.option('-e', '--eventalign-fn', required=False, default='-', show_default=True,
              help=('File containing output of nanopolish eventalign. '
                    'Can be gzipped. Use - to read from stdin'))
@click.option('-s', '--summary-fn', required=False, default=None, show_default=True,
              help=('Summary file containing nanopolish eventalign read names. '
                    'Can be gzipped. Only required if nanopolish was not run '
                    'using --print-read-names options'))
@click.option('-h', '--hdf5-fn', required=True,
              help='Output file (packed event container read by gmmtest)')
@click.option('-g', '--gtf-fn', required=False, default=None,
              help=('Optional GTF file which can be used to convert '
                    'transcriptomic coordinates to genomic'))
@click.option('-p', '--processes', required=False, default=1, show_default=True)
@click.option('-n', '--chunksize', required=False, default=1_000_000, hidden=True)
@click.option('--dense-fn', required=False, default=None,
              help=('Also write the dense sample container gmmtest ingests fastest (the `pack` step done here; '
                    'not in the reference)'))
def nanopolish_collapse(**kw):
    '''
    Parse nanopolish eventalign tabular data into the per-gene event tables
    gmmtest reads (reference prep.py:168-190).
    '''
    opts = CollapseOpts(**kw)
    try:
        import pyarrow.csv  # noqa: F401 -- the table parser of this command
    except ImportError as exc:
        raise click.ClickException('`prep` parses the eventalign table with pyarrow.csv: install pyarrow') from exc
    save_events_to_packed(
        parallel_collapse(opts.eventalign_fn, opts.summary_fn, opts.gtf_fn, opts.processes, opts.chunksize),
        opts.hdf5_fn)
    if getattr(opts, 'dense_fn', None):
        from . import dense
        dense.convert(opts.hdf5_fn, opts.dense_fn, threads=max(1, int(opts.processes)))


if __name__ == '__main__':
    nanopolish_collapse()
