"""
Restriction-site counts per sequence of a multi-FASTA: the part of ContactMap.__init__ (contact_map.py:520-531,
seq_utils.py:95-161 SiteCounter) that precedes the hot path.  FASTA digestion is OUTSIDE the accelerated path (SURVEY.md
section 2, component 9); this minimal host counter exists only so that `ContactMap(bam_path, enzymes, fasta_path, ...)`
-- the call bin3C.py mkmap makes (bin3C.py:148-158) -- works end to end without Biopython.  The reference counts with
Bio.Restriction (`len(enzyme.search(seq, linear=True))` summed over the enzymes, seq_utils.py:134-150): the number of
occurrences of each enzyme's recognition site, overlapping ones included.  Not pinned against Biopython (absent here);
a caller that has its own counts passes them instead (ContactMap's `seq_file` also takes an array, dict or callable).
"""
import gzip
import re
from difflib import SequenceMatcher

from .exceptions import UnknownEnzymeException

# recognition sites (NEB spelling) of the enzymes Hi-C / Meta3C protocols use; IUPAC codes allowed
RECOGNITION = {
    'MluCI': 'AATT', 'Sau3AI': 'GATC', 'DpnII': 'GATC', 'MboI': 'GATC', 'BfuCI': 'GATC', 'HindIII': 'AAGCTT',
    'NcoI': 'CCATGG', 'NlaIII': 'CATG', 'HinfI': 'GANTC', 'DdeI': 'CTNAG', 'MseI': 'TTAA', 'Csp6I': 'GTAC',
    'CviQI': 'GTAC', 'CviAII': 'CATG', 'EcoRI': 'GAATTC', 'BglII': 'AGATCT', 'XhoI': 'CTCGAG', 'SphI': 'GCATGC',
    'ApoI': 'RAATTY', 'HpyCH4IV': 'ACGT', 'AluI': 'AGCT', 'BamHI': 'GGATCC', 'SacI': 'GAGCTC', 'PstI': 'CTGCAG',
    'HpaII': 'CCGG', 'MspI': 'CCGG', 'TaqI': 'TCGA', 'XbaI': 'TCTAGA', 'NheI': 'GCTAGC', 'SalI': 'GTCGAC',
}
_IUPAC = {'A': 'A', 'C': 'C', 'G': 'G', 'T': 'T', 'R': '[AG]', 'Y': '[CT]', 'S': '[CG]', 'W': '[AT]', 'K': '[GT]',
          'M': '[AC]', 'B': '[CGT]', 'D': '[AGT]', 'H': '[ACT]', 'V': '[ACG]', 'N': '[ACGT]'}
_COMP = str.maketrans('ACGTRYSWKMBDHVN', 'TGCAYRSWMKVHDBN')


def _patterns(enzyme_names):
    if isinstance(enzyme_names, str):
        enzyme_names = [enzyme_names]
    pats = []
    for name in enzyme_names:
        site = RECOGNITION.get(name)
        if site is None:        # strict about names, helpful about near misses (seq_utils.py:119-131)
            similar = [k for k in RECOGNITION if SequenceMatcher(None, name.lower(), k.lower()).ratio() >= 0.8]
            raise UnknownEnzymeException(name, similar)
        both = {site, site.translate(_COMP)[::-1]}           # a non-palindromic site is searched on both strands
        pats.extend(re.compile('(?=' + ''.join(_IUPAC[c] for c in s) + ')') for s in sorted(both))
    return pats


def count_sites(seq, patterns):
    """Occurrences of every pattern (overlapping ones included) in an upper-case sequence string."""
    return sum(sum(1 for _ in p.finditer(seq)) for p in patterns)


def tip_sites(seq, patterns, tip_size):
    """[head sites, tail sites] of a tip-based map (SiteCounter.count_sites, seq_utils.py:146-158): the first and last
    tip_size bases, or the two halves of a sequence shorter than 2 tip_size -- with Python 2's integer division,
    seq[:n/2] and seq[-n/2:]: -n/2 rounds towards minus infinity, so the tail half of an odd-length sequence is the
    longer one."""
    n = len(seq)
    if n < 2 * tip_size:
        l_tip, r_tip = seq[:n // 2], seq[(-n) // 2:]
    else:
        l_tip, r_tip = seq[:tip_size], seq[-tip_size:]
    return [count_sites(l_tip, patterns), count_sites(r_tip, patterns)]


def fasta_site_table(path, enzyme_names, min_len=0, tip_size=None):
    """{sequence id: {'sites': n, 'length': L}} for the sequences of at least min_len bases -- the reference's
    `fasta_info` (contact_map.py:520-531).  Plain or gzip'd FASTA.  With tip_size, 'sites' is [head, tail]."""
    pats = _patterns(enzyme_names)
    opener = gzip.open if str(path).endswith('.gz') else open
    info = {}

    def flush(name, parts):
        if name is None:
            return
        seq = ''.join(parts).upper()
        if len(seq) >= min_len:
            info[name] = {'sites': tip_sites(seq, pats, tip_size) if tip_size else count_sites(seq, pats),
                          'length': len(seq)}

    name, parts = None, []
    with opener(path, 'rt') as fh:
        for line in fh:
            if line.startswith('>'):
                flush(name, parts)
                name, parts = line[1:].split()[0] if len(line) > 1 and line[1:].split() else '', []
            else:
                parts.append(line.strip())
    flush(name, parts)
    return info
