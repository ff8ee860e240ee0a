"""
Test infrastructure: a minimal BAM writer (SAM spec section 4: BGZF container, header, alignment records)
so that the host IO library can be exercised without pysam/samtools.  Alignments are dicts with the keys
name, flag, tid, pos, mapq, cigar ([(op, len), ...], op 0 = M, 4 = S ...), and optionally seq_len.
"""
import struct
import zlib


def bgzf_block(data, level=6):
    assert len(data) <= 65280
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    payload = co.compress(data) + co.flush()
    bsize = len(payload) + 25
    head = struct.pack('<BBBBIBBHBBHH', 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord('B'), ord('C'), 2, bsize)
    return head + payload + struct.pack('<II', zlib.crc32(data) & 0xffffffff, len(data))


BGZF_EOF = bgzf_block(b'')


def encode_alignment(a):
    name = a['name'].encode('ascii') + b'\0'
    cigar = a.get('cigar') or []
    l_seq = a.get('seq_len', sum(n for op, n in cigar if op in (0, 1, 4, 7, 8)))
    body = struct.pack('<iiBBHHHiiii', a['tid'], a['pos'], len(name), a['mapq'], 4680, len(cigar), a['flag'], l_seq,
                       a.get('next_tid', -1), a.get('next_pos', -1), a.get('tlen', 0))
    body += name
    body += b''.join(struct.pack('<I', (n << 4) | op) for op, n in cigar)
    body += b'\x11' * ((l_seq + 1) // 2) + b'\x1e' * l_seq
    body += a.get('tags', b'')
    return struct.pack('<i', len(body)) + body


def encode_header(references, lengths, sort_order='queryname', hd=True):
    text = ''
    if hd:
        text += '@HD\tVN:1.6' + ('\tSO:{}'.format(sort_order) if sort_order else '') + '\n'
    for n, l in zip(references, lengths):
        text += '@SQ\tSN:{}\tLN:{}\n'.format(n, l)
    t = text.encode('ascii')
    out = b'BAM\1' + struct.pack('<i', len(t)) + t + struct.pack('<i', len(references))
    for n, l in zip(references, lengths):
        nb = n.encode('ascii') + b'\0'
        out += struct.pack('<i', len(nb)) + nb + struct.pack('<i', l)
    return out


def write_bam(path, references, lengths, alignments, sort_order='queryname', hd=True, block_bytes=65280, level=6,
              eof_marker=True):
    """Write a BAM file; `block_bytes` sets how the byte stream is cut into BGZF blocks (records straddle them)."""
    stream = encode_header(references, lengths, sort_order, hd) + b''.join(encode_alignment(a) for a in alignments)
    with open(path, 'wb') as out:
        for o in range(0, len(stream), block_bytes):
            out.write(bgzf_block(stream[o:o + block_bytes], level))
        if eof_marker:
            out.write(BGZF_EOF)
    return len(stream)
