"""CPU tests of the minimal FASTA site counter that lets ContactMap take the FASTA path bin3C.py mkmap passes."""
import gzip

import numpy as np
import pytest

from bin3c_b200 import seq_sites
from bin3c_b200.exceptions import UnknownEnzymeException


def test_counts_overlapping_and_degenerate_sites(tmp_path):
    pats = seq_sites._patterns(['MluCI'])
    assert seq_sites.count_sites('AATTAATT', pats) == 2 and seq_sites.count_sites('AATT' * 3 + 'A', pats) == 3
    assert seq_sites.count_sites('TTAATTAATT', seq_sites._patterns('MluCI')) == 2
    assert seq_sites.count_sites('GAATC GACTC GANTC'.replace(' ', ''), seq_sites._patterns(['HinfI'])) == 2
    assert seq_sites.count_sites('GATCGATC', seq_sites._patterns(['Sau3AI', 'MluCI'])) == 2
    with pytest.raises(UnknownEnzymeException) as ei:
        seq_sites._patterns(['HindII'])
    assert 'HindIII' in str(ei.value)


@pytest.mark.parametrize('gz', [False, True])
def test_fasta_site_table(tmp_path, gz):
    rng = np.random.default_rng(5)
    seqs = {'a': ''.join(rng.choice(list('ACGT'), 3000)), 'b': 'acgt' * 100, 'c': 'AATT' * 400}
    path = str(tmp_path / ('x.fa.gz' if gz else 'x.fa'))
    with (gzip.open(path, 'wt') if gz else open(path, 'w')) as fh:
        for k, s in seqs.items():
            fh.write('>{} desc\n'.format(k))
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + '\n')
    info = seq_sites.fasta_site_table(path, ['MluCI'], min_len=1000)
    assert sorted(info) == ['a', 'c']                               # b is shorter than min_len
    assert info['c'] == {'sites': 400, 'length': 1600}
    assert info['a']['sites'] == sum(1 for i in range(2997) if seqs['a'][i:i + 4] == 'AATT')
