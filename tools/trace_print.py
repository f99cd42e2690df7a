"""Pretty-print a MONOPORT_B200_TC_TRACE dump (grep "tc trace" lines): one steady-state tile of CTA 0, program v3."""
import sys
rows = [l.split() for l in open(sys.argv[1]) if "tc trace" in l]
d = {}
for r in rows:
    d[int(r[2])] = int(r[3])
names = {0: 'M tile start', 9: 'M L1hid issued', 10: 'M xready', 11: 'M L1skip issued', 12: 'M h1ready', 13: 'M L2 issued',
         14: 'M L3skip issued', 15: 'M h2ready', 16: 'M L3 issued', 96: 'S start', 97: 'S xfree', 98: 'S done'}
for c in range(8):
    names[1 + c] = 'M h0ready c%d' % c
for wg in (0, 1):
    tb = 32 + wg * 32
    for s in range(11):
        names[tb + s] = 'W%d step %d' % (wg, s)
    for k, nm in zip(range(11, 17), ('acc1full', 'drain1 done', 'acc2full', 'drain2 done', 'acc3full', 'drain3 done')):
        names[tb + k] = 'W%d %s' % (wg, nm)
for wg in (0, 1):
    for k, nm in enumerate(('gen4 enter', 'gen4 h0free ok', 'gen4 b0 loads issued', 'gen4 b0 stored', 'gen4 b1 loads issued', 'gen4 b1 stored', 'gen4 arrived')):
        names[32 + wg * 32 + 18 + k] = 'W%d   %s' % (wg, nm)
for k, v in sorted(d.items(), key=lambda kv: kv[1]):
    print("%8d  %s" % (v, names.get(k, k)))
