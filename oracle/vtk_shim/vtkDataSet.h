/* stand-in for VTK's vtkDataSet.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
