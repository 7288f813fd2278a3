!  sizes for tests/f03c_cases/semantics.f03 (the translator expects the five run-time sizes npc, mx, my, mz, np0)
      integer(C_INT) npc,mx,my,mz,np0,nn
      parameter  (npc=2)
      parameter  (mx=4,my=3,mz=4)
      parameter  (np0=10)
      parameter  (nn=mx*my)
