"""Hand-traced golden vectors: /root/reference/src/dbg.jl executed BY HAND, line by line, on inputs
small enough to do so (k = 3 / 4).  They are independent of tests/pyref.py and of oracle/ -- the
traces below were written from the Julia source alone and only afterwards compared with both.
They are still not outputs of the reference run by Julia (there is none in this image);
oracle/make_ref_golden.jl regenerates every vector of this directory with the real package.

Conventions used in the traces (reference lines in brackets): list positions are 1-based;
`c(X)` = canonical(X) = min(X, revcomp(X)); get_fw_idxs!/get_bw_idxs! [:70-90] look the canonical
form of each of the four oriented neighbours up and keep the ORIENTED neighbour + list position;
is_end_fw [:57-68]: not exactly one fw neighbour, or that neighbour has not exactly one bw
neighbour; is_end_bw [:44-55] likewise; add_link!(s, d, dist) [src/Graphs.jl:473-476] pushes
(s, d, dist) to links[|s|] and then (d, s, dist) to links[|d|].

G1 (traced by the surveyor): docs/src/DeBruijnGraph.md:57-58 canonicalised.

G2 -- hairpin stop (`used_kmers` break, src/dbg.jl:147-150), k = 3, list [ACG, AGA, GAC, TCA]
  1 ACG: bw nbrs {AAC, CAC, GAC*, TAC->GTA} -> [GAC]; fw of GAC {ACA, ACC, ACG*, ACT} -> 1  => not end_bw
         fw nbrs {CGA, CGC, CGG->CCG, CGT->ACG*} -> [CGT@1]; bw of CGT {ACG*, CCG, GCG->CGC, TCG->CGA} -> 1 => not end_fw
         middle of a unitig [:119-122] -> skipped
  2 AGA: bw nbrs {AAG, CAG, GAG->CTC, TAG->CTA} -> 0 => end_bw; fw nbrs {GAA, GAC*, GAG, GAT->ATC} -> [GAC];
         bw of GAC {AGA*, CGA, GGA, TGA->TCA*} -> 2 => end_fw.  Single k-mer [:124-128]: node 1 = c!(AGA) = AGA
  3 GAC: bw nbrs -> [AGA, TGA] (2) => end_bw; fw nbrs -> [ACG]; bw of ACG -> [GAC] (1) => not end_fw.
         unitig starts forwards [:130-140]: s = GAC; fw -> ACG@1 unused -> s = GACG, used[1]; is_end_fw(ACG) = false;
         fw of ACG -> CGT@1, used[1] already set -> break [:147-150].  node 2 = c!(GACG) = min(GACG, CGTC) = CGTC
  4 TCA: bw nbrs {ATC, CTC, GTC->GAC*, TTC->GAA} -> [GTC@3]; fw of GTC {TCA*, TCC->GGA, TCG->CGA, TCT->AGA*} -> 2 => end_bw;
         fw nbrs {CAA, CAC, CAG, CAT->ATG} -> 0 => end_fw.  node 3 = c!(TCA) = TCA
  no unused k-mer left -> no circle pass.
  overlaps [:199-218]: node 1 AGA: first AG canonical -> in (AG, 1); last GA canonical -> out (GA, -1)
                       node 2 CGTC: first CG (palindrome, iscanonical) -> in (CG, 2); last TC not canonical -> in (GA, -2)
                       node 3 TCA: first TC not canonical -> out (GA, 3); last CA canonical -> out (CA, -3)
  in = [(AG,1), (CG,2), (GA,-2)], out = [(CA,-3), (GA,-1), (GA,3)]
  merge [:229-239]: (GA,-2) meets (GA,-1), (GA,3) -> add_link!(-2,-1,-2), add_link!(-2,3,-2)

G3 -- palindromic (k-1)-mers go to `in` at a node start and to `out` at a node end only (src/dbg.jl:203-217),
  k = 3, list [AAT, ATC, ATG]
  1 AAT: bw nbrs {AAA, CAA, GAA, TAA} -> 0 => end_bw; fw nbrs {ATA, ATC*, ATG*, ATT->AAT*} -> 3 => end_fw.  node 1 = AAT
  2 ATC: bw nbrs {AAT*, CAT->ATG*, GAT->ATC*, TAT->ATA} -> 3 => end_bw; fw nbrs {TCA, TCC->GGA, TCG->CGA, TCT->AGA} -> 0 => end_fw.  node 2 = ATC
  3 ATG: bw nbrs {AAT*, CAT->ATG*, GAT->ATC*, TAT} -> 3 => end_bw; fw nbrs {TGA->TCA, TGC->GCA, TGG->CCA, TGT->ACA} -> 0 => end_fw.  node 3 = ATG
  overlaps: node 1 AAT: in (AA, 1), last AT palindrome -> out (AT, -1); node 2 ATC: first AT -> in (AT, 2), last TC -> in (GA, -2);
            node 3 ATG: in (AT, 3), last TG -> in (CA, -3)
  in = [(AA,1), (AT,2), (AT,3), (CA,-3), (GA,-2)], out = [(AT,-1)]
  merge: (AT,2) x (AT,-1) -> add_link!(2,-1,-2); (AT,3) x (AT,-1) -> add_link!(3,-1,-2)   (next_out_idx does not move past an equal mer)
  NOT created, by construction of in/out: start of 2 <-> start of 3 (GAT -> ATG), end of 1 <-> end of 1 (AAT -> ATT).

G4 -- perfect circle pass (src/dbg.jl:160-181), paths numbered before circles, k = 3, list [AAC, ACA, AGG, CAA, CTC]
  1 AAC: fw -> [ACA]; bw of ACA -> [AAC] => not end_fw; bw -> [CAA]; fw of CAA -> [AAC] => not end_bw -> skipped
  2 ACA: fw -> [CAA], bw of CAA -> [ACA]; bw -> [AAC], fw of AAC -> [ACA] -> skipped
  3 AGG: bw nbrs {AAG, CAG, GAG->CTC*, TAG->CTA} -> [GAG@5]; fw of GAG {AGA, AGC, AGG*, AGT->ACT} -> 1 => not end_bw;
         fw nbrs {GGA, GGC->GCC, GGG->CCC, GGT->ACC} -> 0 => end_fw.  The unitig must start forwards [:134-137]:
         current = revcomp(AGG) = CCT, end_fw = end_bw = false; s = CCT; fw of CCT {CTA, CTC*, CTG->CAG, CTT->AAG} -> CTC@5 unused,
         s = CCTC; is_end_fw(CTC): fw nbrs {TCA, TCC->GGA, TCG->CGA, TCT->AGA} -> 0 => true.  node 1 = c!(CCTC) = min(CCTC, GAGG) = CCTC
  4 CAA: fw -> [AAC], bw of AAC -> [CAA]; bw -> [ACA], fw of ACA -> [CAA] -> skipped
  5 CTC: used
  circle pass: next_unused = 1 (AAC): s = AAC; fw -> ACA@2, s = AACA; fw -> CAA@4, s = AACAA; fw -> AAC == start -> break.
         node 2 = c!(AACAA) = min(AACAA, TTGTT) = AACAA
  overlaps: node 1 CCTC: in (CC, 1), last TC -> in (GA, -1); node 2 AACAA: in (AA, 2), last AA canonical -> out (AA, -2)
  merge: (AA,2) x (AA,-2) -> add_link!(2,-2,-2): both pushes land in links[2]

T1 / T2 -- tip removal: src/Graphs.jl (find_tip_nodes! :778-802, find_all_unitigs! :831-862, join_path! :1040-1088, remove_node!
  :459-466, remove_link! :487-495) and src/remove_tips.jl:2-30 executed by hand.  forward_links(n) = entries of links[|n|] with
  source == -n, backward_links(n) = forward_links(-n) [:667-690].

T1 (fork with a tip), nodes 1 AAAC, 2 AACGGG, 3 AACT, 4 CGGGT, 5 TTTTT; add_link!(-1,2,-3), (-1,3,-3), (-2,4,-4):
  links[1] = [(-1,2), (-1,3)], links[2] = [(2,-1), (-2,4)], links[3] = [(3,-1)], links[4] = [(4,-2)], links[5] = []
  find_tip_nodes(4): node 1: fw 2 links -> no rule applies; node 2 (6 bases > 4) skipped; node 3: fw 0, bw 1 (to -1),
  forward_links(1) has 2 > 1 -> tip; node 4 (5 bases) skipped; node 5 (5 bases) skipped.  tips = {3}.
  remove_node!(3): links[3] = [], links[1] = [(-1,2)].  collapse_all_unitigs!(.., 2, true): n = 1: fw [(-1,2)], dest 2,
  backward_links(2) = [(2,-1)] (1) -> push 2; fw(2) = [(-2,4)], backward_links(4) = [(4,-2)] -> push 4; fw(4) = []; reversed
  twice -> path [1, 2, 4] (node 5: alone).  sequence: AAAC + AACGGG[4:] (overlap AAC ok) = AAACGGG, + CGGGT[5:] (overlap CGGG
  ok) = AAACGGGT; canonical (A < comp(T) = A? equal; A vs comp(G) = C: A < C) -> kept.  node 6 = AAACGGGT without links (first has
  no backward, last no forward link); nodes 1, 2, 4 emptied.  Pass 2: node 6 has 8 bases > 4, node 5 has 5 > 4: no tip.

T2 (two joined paths next to a hub: link ORDER), nodes 1 AC, 2 AAA, 3 C, 4 GG, 5 T, all distances 1 (no overlap);
  add_link!(-1,2), (-1,4), (-2,3), (-4,5), (-3,-5), (-3,1):
  links[1] = [(-1,2), (-1,4), (1,-3)]   links[2] = [(2,-1), (-2,3)]   links[3] = [(3,-2), (-3,-5), (-3,1)]
  links[4] = [(4,-1), (-4,5)]           links[5] = [(5,-4), (-5,-3)]
  find_all_unitigs!(2): n = 1: fw has 2 links; backwards: fw(-1) = [(1,-3)], backward_links(-3) = fw(3) has 2 -> stop: [1] alone.
    n = 2: fw [(-2,3)], backward_links(3) = [(3,-2)] -> push 3; fw(3) has 2 -> stop; backwards: fw(-2) = [(2,-1)], node 1 consumed
    -> stop: [2, 3].  n = 4: fw [(-4,5)], backward_links(5) = [(5,-4)] -> push 5; fw(5) = [(-5,-3)], node 3 consumed -> stop;
    backwards: fw(-4) = [(4,-1)], node 1 consumed: [4, 5].
  join [2, 3]: AAA + C = AAAC canonical (A < comp(C) = G); node 6.  backward_links(2) = [(2,-1)] -> add_link!(6,-1);
    forward_links(3) = [(-3,-5), (-3,1)] -> add_link!(-6,-5), add_link!(-6,1).  Nodes 2, 3 removed with their entries:
    links[1] = [(-1,4), (-1,6), (1,-6)]   links[4] = [(4,-1), (-4,5)]   links[5] = [(5,-4), (-5,-6)]
    links[6] = [(6,-1), (-6,-5), (-6,1)]
  join [4, 5]: GG + T = GGT, not canonical (comp(T) = A < G) -> ACC, path becomes [-5, -4]; node 7.
    backward_links(-5) = entries of links[5] with source -5 = [(-5,-6)] -> add_link!(7,-6);
    forward_links(-4) = entries of links[4] with source 4 = [(4,-1)] -> add_link!(-7,-1).  Nodes 4, 5 removed:
    links[1] = [(-1,6), (1,-6), (-1,-7)]   links[6] = [(6,-1), (-6,1), (-6,7)]   links[7] = [(7,-6), (-7,-1)]
"""

HAND_TRACED = [
    dict(name="G1_doc_example", k=4, kmers=["AATC", "AATG", "ACGA", "ATTC", "CGAA", "CGTA", "CGTC"],
         nodes=["AATC", "AATG", "ACGAAT", "CGTA", "CGTC"],
         links=[[(1, -3, -3)], [(2, -3, -3)], [(-3, 1, -3), (-3, 2, -3), (3, 4, -3), (3, 5, -3)], [(4, 3, -3)], [(5, 3, -3)]]),
    dict(name="G2_hairpin", k=3, kmers=["ACG", "AGA", "GAC", "TCA"],
         nodes=["AGA", "CGTC", "TCA"],
         links=[[(-1, -2, -2)], [(-2, -1, -2), (-2, 3, -2)], [(3, -2, -2)]]),
    dict(name="G3_palindromic_overlap", k=3, kmers=["AAT", "ATC", "ATG"],
         nodes=["AAT", "ATC", "ATG"],
         links=[[(-1, 2, -2), (-1, 3, -2)], [(2, -1, -2)], [(3, -1, -2)]]),
    dict(name="G4_circle_after_paths", k=3, kmers=["AAC", "ACA", "AGG", "CAA", "CTC"],
         nodes=["CCTC", "AACAA"],
         links=[[], [(2, -2, -2), (-2, 2, -2)]]),
]

# graph editing (see T1 / T2 above): input graph, min_size for remove_tips!, expected graph after
# collapse_all_unitigs!(sg, 2, true) on the input ("collapse") and after remove_tips!(sg, min_size) ("remove_tips")
TIPS_HAND_TRACED = [
    dict(name="T1_fork_with_a_tip", min_size=4,
         nodes=["AAAC", "AACGGG", "AACT", "CGGGT", "TTTTT"],
         links=[[(-1, 2, -3), (-1, 3, -3)], [(2, -1, -3), (-2, 4, -4)], [(3, -1, -3)], [(4, -2, -4)], []],
         tips=[3],
         remove_tips=dict(passes=2, removed=1, nodes=["", "", "", "", "TTTTT", "AAACGGGT"], links=[[], [], [], [], [], []])),
    dict(name="T2_two_paths_next_to_a_hub", min_size=0,
         nodes=["AC", "AAA", "C", "GG", "T"],
         links=[[(-1, 2, 1), (-1, 4, 1), (1, -3, 1)], [(2, -1, 1), (-2, 3, 1)], [(3, -2, 1), (-3, -5, 1), (-3, 1, 1)],
                [(4, -1, 1), (-4, 5, 1)], [(5, -4, 1), (-5, -3, 1)]],
         tips=[],
         collapse=dict(nodes=["AC", "", "", "", "", "AAAC", "ACC"],
                       links=[[(-1, 6, 1), (1, -6, 1), (-1, -7, 1)], [], [], [], [], [(6, -1, 1), (-6, 1, 1), (-6, 7, 1)],
                              [(7, -6, 1), (-7, -1, 1)]])),
]
