#include "freesasa_b200_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
freesasa_structure *freesasa_structure_from_pdb_buffer(const char *text, long len, const freesasa_classifier *c, int options);
int main(int argc,char**argv){
  freesasa_set_verbosity(FREESASA_V_SILENT);
  for (int a=1;a<argc;a++){
    FILE*f=fopen(argv[a],"r"); if(!f) continue;
    for (int opt=0; opt<512; opt+= 37) {
      rewind(f);
      freesasa_structure*s=freesasa_structure_from_pdb(f,NULL,opt & ~(8|16));
      if(s){ int n=freesasa_structure_n(s); freesasa_result r; r.n_atoms=n; r.sasa=malloc(8*n); for(int i=0;i<n;i++) r.sasa[i]=i%7; r.total=1; r.parameters=freesasa_default_parameters;
        freesasa_node*t=freesasa_tree_init(&r,s,"x"); freesasa_node *t2=freesasa_tree_init(&r,s,NULL); freesasa_tree_join(t,&t2);
        FILE*o=fopen("/dev/null","w"); freesasa_write_pdb(o,t); fclose(o);
        { const char *cmds[] = {"a, resn ALA+GLY and not chain B", "b, resi 1-20+\\-3 or name CA", "c, resn ALAA", "d resn", "e, resn ALA and"};
          for (int q=0;q<5;q++){ freesasa_selection *sel=freesasa_selection_new(cmds[q],s,&r); if(sel){ freesasa_selection *cl=freesasa_selection_clone(sel);
            freesasa_node_structure_add_selection(freesasa_node_children(freesasa_node_children(t)), sel); freesasa_selection_free(cl); freesasa_selection_free(sel);} } }
        freesasa_structure*c=freesasa_structure_get_chains(s,"A",NULL,0); freesasa_structure_free(c);
        (void)freesasa_structure_atom_pdb_line(s,0);
        freesasa_node_free(t); free(r.sasa); freesasa_structure_free(s);} 
      int cnt=0; rewind(f); freesasa_structure**arr=freesasa_structure_array(f,&cnt,NULL,8|16|(opt&5)); if(arr){for(int i=0;i<cnt;i++)freesasa_structure_free(arr[i]); free(arr);} 
    }
    fclose(f);
  }
  const char *cfg="name: t\ntypes:\nA 1.0 polar\nB 2 apolar\natoms:\nAA aa A\nBB bb B\n"; FILE*m=fmemopen((void*)cfg,strlen(cfg),"r"); freesasa_classifier*c=freesasa_classifier_from_file(m); fclose(m); freesasa_classifier_free(c);
  const char *bad="name: t\ntypes:\nA 1.0 polar\natoms:\nAA aa Q\n"; m=fmemopen((void*)bad,strlen(bad),"r"); c=freesasa_classifier_from_file(m); fclose(m); freesasa_classifier_free(c);
  puts("done"); return 0; }
