"""TEST INFRASTRUCTURE ONLY.  CPU restatements of the reference's hot path; the product
package (biceps_b200/) must never import anything from here."""
